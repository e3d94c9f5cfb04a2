/*
 * vqgnn.h — C-ABI of libvqgnn.so: the B200 (sm_100a) kernels behind VQ-GNN's hot path.
 *
 * The reference (devnkong/VQ-GNN) is pure Python and has no FFI; its boundary for this path is the
 * Python class surface.  Each entry point below replaces the body of one reference
 * function; the Python modules in vq_gnn_b200/ keep the reference's class names and signatures and
 * bind these symbols with ctypes (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named h_*;
 *   - no allocation, no ownership transfer: the caller passes outputs and workspaces;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), stateless, thread-safe;
 *   - return 0 on success, a negative VQGNN_ERR_* otherwise; vqgnn_last_error() gives the text;
 *   - all floating point is fp32 (moments are accumulated in fp64), indices int32, codes int16.
 *
 * Layouts (HBM)
 *   x / g / y      row-major [rows, ld] fp32; branch k owns columns [k*D, (k+1)*D) of x and
 *                  [k*Dg, (k+1)*Dg) of g (Dg = D, or D+1 for the v1 GAT "add_flag" quantiser);
 *   codebooks      E (whitened, `_embedding`), W (`_ema_w`), O (de-whitened, `_embedding_output`):
 *                  [nb, M, Wp] with Wp = 4*ceil(W/4), W = D + Dg (feature part first, gradient part after);
 *   sizes          `_ema_cluster_size` [nb, M];
 *   codes          node -> codeword table [N, nb] int16 (one row = all branches of a node; the
 *                  reference keeps nb separate `c_indices[N]` buffers, vq_gnn_v2/models.py:27-28);
 *   stats          per-codeword accumulator [nb, M, Wp + 4]: Wp sums, then count, then 3 pad words;
 *                  this flat buffer (plus `sums`) is what a multi-GPU run allreduces between
 *                  vqgnn_vq_segsum and vqgnn_vq_finalize.
 *
 * Determinism: results are pure functions of the inputs -- no entry point of the VQ update, the plan builders or
 * the GCN / SAGE message passing accumulates floating point with atomics whose order could change the sum
 * (two-level ordered reductions instead; integer counters only).  The GAT kernels still use fp32 REDs.
 */
#ifndef VQGNN_H_
#define VQGNN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VQGNN_ABI_VERSION 3

#define VQGNN_OK 0
#define VQGNN_ERR_ARCH (-1)      /* device is not sm_100 */
#define VQGNN_ERR_ARG (-2)       /* bad argument / unsupported shape */
#define VQGNN_ERR_CUDA (-3)      /* a CUDA runtime call failed */
#define VQGNN_ERR_WORKSPACE (-4) /* workspace too small */

/* status word bits written by vqgnn_vq_finalize (read lazily by the host) */
#define VQGNN_STATUS_BAD_INIT 1  /* some EMA cluster size is exactly 0: reference raises ValueError('Bad Init!') (vq.py:188,253) */

int vqgnn_abi_version(void);
/* 0 if `device` is compute capability 10.x, VQGNN_ERR_ARCH otherwise. */
int vqgnn_arch_check(int device);
const char* vqgnn_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches claim) */
int64_t vqgnn_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * VectorQuantizerEMA  (vq_gnn_v2/vq.py:160-279), all nb branches of a layer in one launch each.
 * ------------------------------------------------------------------------------------------- */

/* Column sums and sums of squares (fp64) of x[B, C] and, if g != NULL, g[B, Cg]:
 *   sums[0 .. C+Cg)          = sum_b v[b, c]
 *   sums[C+Cg .. 2(C+Cg))    = sum_b v[b, c]^2
 * Replaces the batch statistics inside BatchNorm1d.forward (vq.py:162,223) and the explicit
 * mean/var at vq.py:216-221.  `sums` is overwritten.  In a multi-GPU run it is allreduced (sum).
 * Two-level ordered reduction (row tiles of 512, then the tiles in index order): bit-stable.
 * ws: vqgnn_vq_moments_workspace_bytes(B, C, Cg) bytes, 8 B aligned. */
size_t vqgnn_vq_moments_workspace_bytes(int64_t B, int C, int Cg);
int vqgnn_vq_moments(const float* x, int64_t ldx, const float* g, int64_t ldg, int64_t B, int C, int Cg,
                     double* sums, void* ws, void* stream);

/* Turns moments into the per-column whitening affine z = v * scale + shift and updates the BatchNorm
 * running statistics exactly as torch.nn.BatchNorm1d(affine=False) does in train mode
 * (biased variance for normalisation, unbiased for the running update), incl. the first-update
 * re-seeding of vq.py:216-221 (`seed_running` != 0, or per branch seed_mask[k] != 0 when seed_mask != NULL:
 * the reference keeps one `bn_inited` flag per quantiser).  If training == 0 the running statistics are
 * used and left untouched, except that a seeding branch first overwrites them with the batch statistics.  Gradient columns are additionally multiplied by grad_scale0 (last column
 * of each branch by grad_scale1 when Dg == D+1) (vq.py:224-227).
 *   run_mean_f/run_var_f: [nb*D]; run_mean_g/run_var_g: [nb*Dg] (NULL when Cg == 0)
 *   scale/shift: [C + Cg] outputs.  count = number of rows the moments were taken over (global); if d_count
 *   (a device double) is not NULL it overrides `count`, so an allreduced row count needs no host sync.
 *   nbt_f [nb] / nbt_g [nb]: BatchNorm1d.num_batches_tracked counters (int64, may be NULL), +1 in training. */
int vqgnn_vq_whiten(const double* sums, double count, const double* d_count, int nb, int D, int Dg, int has_grad,
                    float* run_mean_f, float* run_var_f, float* run_mean_g, float* run_var_g,
                    float eps_f, float mom_f, float eps_g, float mom_g, float grad_scale0, float grad_scale1,
                    int training, int seed_running, const int32_t* seed_mask, int64_t* nbt_f, int64_t* nbt_g,
                    float* scale, float* shift, void* stream);

/* Fused whitening + nearest-codeword assignment + per-codeword accumulation.
 * For every row b and branch k: z = whiten([x_k | g_k]); code = argmin_m ||z||^2 + ||E_k[m]||^2 - 2 z.E_k[m]
 * (that association order, fp32, lowest index wins ties: vq.py:166-171 / 230-236) using the
 * PRE-update codebook; writes idx[b, k] ([B, nb]) and, if codes != NULL, codes[batch_idx[b]*codes_ld + k]
 * (models.py:46,63; codes_ld = row stride of the code table, so a sub-range of branches can be updated).
 * stats: normally NULL -- the per-codeword sums are produced afterwards by vqgnn_vq_segsum (ordered, bit-stable).
 * If stats != NULL the kernel instead adds z to stats[k, code, :W] and 1 to stats[k, code, Wp] with fp32 atomics in
 * its epilogue (order-dependent last bits; `stats` must be zeroed by the caller with vqgnn_fill_zero).  g == NULL selects the feature-only form (feature_update, W = D).
 * impl: 0 = exact-fp32 SIMT kernel (parity anchor; ws unused, may be NULL);
 *       1 = tcgen05 / TMEM kernel (kind::tf32, error-compensated 3xTF32, TMA-fed codebook tiles; packed width
 *           D+Dg in {4, 8, 9}); needs ws of vqgnn_vq_assign_workspace_bytes() bytes (the codebook re-packed into MMA
 *           tiles), 16 B aligned, a 16 B aligned codebook and idx != NULL.  Two kernels: the epilogue of the
 *           tensor-core kernel keeps, per row, the running minimum over chunks of 8 codewords and the chunk it came from
 *           (written to idx); a second kernel re-scores that chunk in fp32 from E and picks the lowest-index minimum.
 *           Codes may differ from impl 0 only at near-ties of the fp32 distances. */
size_t vqgnn_vq_assign_workspace_bytes(int nb, int M);
int vqgnn_vq_assign(const float* x, int64_t ldx, const float* g, int64_t ldg, const float* scale,
                    const float* shift, const float* E, int64_t B, int nb, int M, int D, int Dg, int Wp,
                    const int32_t* batch_idx, int16_t* codes, int64_t codes_ld, int16_t* idx, float* stats,
                    int impl, void* ws, size_t ws_bytes, void* stream);

/* Per-codeword statistics of the update (the one-hot^T @ z GEMM and the one-hot column sum of vq.py:177,191 /
 * 243,256) as an ordered segmented sum over the assignments idx [B, nbc] written by vqgnn_vq_assign:
 *   stats[k, m, :W] = sum_{b: idx[b,k]==m} z[b, k, :],  stats[k, m, Wp] = count   (every entry is written).
 * The (branch, code) keys are radix-sorted (stable), then one warp per codeword adds its rows in ascending order with
 * a fixed reduction tree: no floating-point atomics, identical bits on every run.  z is re-whitened from x / g with
 * scale / shift exactly as in vqgnn_vq_assign.  ws: vqgnn_vq_segsum_workspace_bytes(B, nbc, M) bytes. */
size_t vqgnn_vq_segsum_workspace_bytes(int64_t B, int nbc, int M);
int vqgnn_vq_segsum(const float* x, int64_t ldx, const float* g, int64_t ldg, const float* scale,
                    const float* shift, const int16_t* idx, int64_t B, int nbc, int M, int D, int Dg, int Wp,
                    float* stats, void* ws, size_t ws_bytes, void* stream);

/* EMA + Laplace smoothing + codeword recovery (vq.py:177-200 / 242-275); sizes by one CTA per branch, then
 * (branch, 256-codeword tile) CTAs for the sums / codewords:
 *   size <- decay*size + (1-decay)*count; if warm_up: size <- (size+1e-5)/(sum(size)+M*1e-5)*sum(size);
 *   status |= BAD_INIT if any size == 0 and, as the reference raises before touching them (vq.py:188,253), Wm / E /
 *   O of that branch are then left unchanged; otherwise Wm <- decay*Wm + (1-decay)*sum_z; E <- Wm/size;
 *   O <- de-whiten(E) with the (already updated) running statistics; gradient part divided by
 *   (grad_scale + eps); O[:, D:] <- 0 if grad_scale0 == 0.
 * joint == 0 updates only the first D columns (feature_update). */
int vqgnn_vq_finalize(const float* stats, int nb, int M, int D, int Dg, int Wp, int joint, double decay,
                      int warm_up, float eps, float grad_scale0, float grad_scale1,
                      const float* run_mean_f, const float* run_var_f, const float* run_mean_g,
                      const float* run_var_g, float* ema_size, float* ema_w, float* E, float* O,
                      int32_t* status, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Message passing, GCN / SAGE-Mean (OurGCNConv.forward = adj @ x, vq_gnn_v2/convs.py:65-101)
 * fused with the codeword gather of vq_gnn_v2/models.py:161-173 (v2) or the A*R product that
 * vq_gnn_v1/utils/dataloader.py:144-192 (`mapper`) materialises (v1).
 * ------------------------------------------------------------------------------------------- */

/* Work partition of a CSR for the message-passing kernels: the entries are cut into chunks of `chunk`
 * consecutive entries (a multiple of 32) regardless of row boundaries, so power-law hub rows are spread
 * over many warps.  chunk_row[c] = the row that entry c*chunk belongs to; vqgnn_mp_num_chunks() entries.
 * Built once per mini-batch (per CSR) and reused by every layer. */
int64_t vqgnn_mp_num_chunks(int64_t nnz, int chunk);
int vqgnn_mp_chunk_rows(const int32_t* rowptr, int64_t R, int64_t nnz, int chunk, int32_t* chunk_row,
                        void* stream);

/* Forward over R rows of the plan's CSR (see vq_gnn_b200/graph.py: BatchPlan).  For row r:
 *   acc   = sum_e val[e] * (col[e] < B ? x[col[e], :] : feat_scale * O_k[code(node(col[e]-B), k), :D])
 *   gqacc = sum_{tail e} rval[e] * O_k[code(...), D:2D]                   (only if rval != NULL)
 * rows r <  B: y[r, :] = acc; gq[r, :] = gqacc; info += <x[r, :], gqacc>  (v1 info_backward, models.py:223)
 * rows r >= B: info += <acc, O_k[code(node(r-B), k), D:2D]>               (v2 info_backward, models.py:198)
 * *info = info_scale * info (fp64; per-block partials added in block order by the last block).
 * C = nb*D columns.  tail_node == NULL means identity.  info/gq may be NULL.
 * chunk_row/chunk/nnz: the partition above for (rowptr, R).
 * Order-independent by construction: a row inside one chunk is stored, a row cut once takes two REDs onto zero
 * (commutative), the pieces of a row spanning >= 3 chunks go to a piece buffer and are added in chunk order.
 * ws: vqgnn_mp_workspace_bytes(nnz, chunk, nb*D) bytes (info partials + piece buffers); same for vqgnn_mp_bwd. */
size_t vqgnn_mp_workspace_bytes(int64_t nnz, int chunk, int C);
int vqgnn_mp_fwd(const int32_t* rowptr, const int32_t* col, const float* val, const float* rval,
                 const int32_t* chunk_row, int chunk, int64_t nnz, int64_t R, int64_t B, const float* x,
                 int64_t ldx, const int32_t* tail_node, const int16_t* codes, const float* O, int nb, int M,
                 int D, int Wp, const float* tail_feat, int64_t ld_tail, int tail_slab, float feat_scale,
                 float info_scale, float* y, int64_t ldy, float* gq, int64_t ldgq, float* info, void* ws,
                 void* stream);

/* The same forward for batch graphs whose out-of-batch codeword rows were materialised (vqgnn_tail_materialize):
 * every entry's row -- x[col] for a batch column, tail_feat[col - B] otherwise -- is fetched by one TMA bulk copy
 * (cp.async.bulk, <= 512 B per 128-column slab) into a per-warp ring of shared memory, so the consuming loop holds no
 * address arithmetic (csrc/mp_rows.cuh).  rows r < B: y[r] = acc; rows r >= B: info += <acc, tail_grad[r - B]>
 * (vq_gnn_v2/models.py:161-198).  Replaces vqgnn_mp_fwd(rval = NULL, tail_slab = 0, feat_scale = 1) with identical
 * per-row arithmetic; needs C >= 16, C % 4 == 0, chunk <= 256, 16 B aligned rows.  tail_feat may be NULL when R == B
 * and every column is < B (plain convolution); tail_grad may be NULL when info is.  ws as for vqgnn_mp_fwd.
 * The value of an entry whose column is >= B is multiplied by tail_scale * (tail_scale_dev ? *tail_scale_dev : 1)
 * (forward: 1, NULL).  That makes the v2 BACKWARD the same call over the transposed CSR of the batch columns:
 * x = dY, tail_feat = the gradient codeword rows, tail_scale = warm-up rate, tail_scale_dev = d info (device scalar),
 * R = B, info = NULL  ==  vqgnn_mp_bwd(gq = NULL) with the same per-element arithmetic.  T = rows of tail_feat. */
int vqgnn_mp_fwd_rows(const int32_t* rowptr, const int32_t* col, const float* val, const int32_t* chunk_row,
                      int chunk, int64_t nnz, int64_t R, int64_t B, const float* x, int64_t ldx,
                      const float* tail_feat, int64_t T, float tail_scale, const float* tail_scale_dev,
                      const float* tail_grad, int64_t ld_tail, int C, float info_scale, float* y, int64_t ldy,
                      float* info, void* ws, void* stream);

/* Dense copies of the tail entries' codewords: tail_feat[t, 4k:4k+4] = O_k[code_k(node(t)), :4],
 * tail_grad[t, 4k:4k+4] = O_k[code_k(node(t)), 4:8] (either may be NULL; D == 4, Wp == 8).  In a v2 batch graph
 * a tail node is referenced by ~20 edges: gathering its codewords once and handing the rows to vqgnn_mp_fwd
 * (tail_feat) / vqgnn_mp_bwd (tail_grad) turns nb scattered sector reads per edge into one coalesced row read.
 * Both kernels fall back to per-edge codebook gathers when the pointer is NULL.  tail_slab > 0: the table handed to
 * vqgnn_mp_fwd / vqgnn_mp_bwd is the SLAB-MAJOR one of vqgnn_tail_materialize_slab ([ceil(C/tail_slab)][T][tail_slab])
 * and ld_tail carries T; tail_slab == 0: row-major [T, ld_tail]. */
int vqgnn_tail_materialize(const int32_t* tail_node, int64_t T, const int16_t* codes, const float* O, int nb,
                           int M, int D, int Wp, float* tail_feat, float* tail_grad, int64_t ld_tail,
                           void* stream);

/* info_backward of the v2 formulation over the out-of-batch rows only (csrc/mp_info.cu): the entries [e_begin, nnz) of
 * the plan's CSR are those of rows r >= B (e_begin = rowptr[B]; erow [nnz] holds their row ids, vqgnn_csr_expand_rows);
 *   *info = info_scale * sum_e val[e] * < Xin[col[e], :], Gq[row(e), :] >          (vq_gnn_v2/models.py:198)
 * with Xin = x for batch columns and the slab-major feature table tfS otherwise, Gq from the slab-major gradient
 * table tgS.  vqgnn_tail_materialize_slab fills tfS / tgS [ceil(C / slab)][T][slab] (zero padded; slab in
 * {16, 32, 64} columns): processed slab by slab the gathered slice stays L2-resident.  Pair it with vqgnn_mp_fwd over
 * the first B rows (R = B, nnz = rowptr[B], info = NULL) for y.  ws: vqgnn_mp_info_workspace_bytes(nnz, C, slab). */
int vqgnn_tail_materialize_slab(const int32_t* tail_node, int64_t T, const int16_t* codes, const float* O, int nb,
                                int M, int D, int Wp, int slab, float* tfS, float* tgS, void* stream);
size_t vqgnn_mp_info_workspace_bytes(int64_t nnz, int C, int slab);
/* erow[e] = row of CSR entry e for the rows [r_begin, R) (the per-entry row ids vqgnn_mp_info streams). */
int vqgnn_csr_expand_rows(const int32_t* rowptr, int64_t r_begin, int64_t R, int32_t* erow, void* stream);
int vqgnn_mp_info(const int32_t* erow, const int32_t* col, const float* val, int64_t e_begin, int64_t nnz,
                  int64_t B, int64_t R, const float* x, int64_t ldx, const float* tfS, const float* tgS, int C,
                  int slab, float info_scale, float* info, void* ws, void* stream);

/* Backward of the same: for batch column j < B
 *   dx[j, :] = sum_e bval[e] * (brow[e] < B ? dy[brow[e], :]
 *                                           : tail_scale * (*dinfo) * O_k[code(node(brow[e]-B), k), D:2D])
 *              + gq_scale * (*dinfo) * gq[j, :]        (if gq != NULL; v1)
 * i.e. adj^T @ dY restricted to batch rows plus the gradient of info_backward.
 * dinfo is a device scalar (NULL = 1).  chunk_row/chunk/nnz: the partition of (browptr, B). */
int vqgnn_mp_bwd(const int32_t* browptr, const int32_t* brow, const float* bval, const int32_t* chunk_row,
                 int chunk, int64_t nnz, int64_t B, const float* dy, int64_t lddy, const int32_t* tail_node,
                 const int16_t* codes, const float* O, int nb, int M, int D, int Wp, const float* tail_grad,
                 int64_t ld_tail, int tail_slab, float tail_scale, const float* gq, int64_t ldgq, float gq_scale,
                 const float* dinfo, float* dx, int64_t lddx, void* ws, void* stream);

/* Out-of-batch ("tail") part of the forward with the codebooks of a branch group resident in shared
 * memory (csrc/mp_tail.cu) -- the fast path for the v1 formulation, where every batch row has hundreds of
 * out-of-batch neighbours (vq_gnn_v1/utils/dataloader.py:149-154 + vq_gnn_v1/models.py:181-223).
 *   vqgnn_mp_tail_group(M, D, Wp): branches per group G (8), or 0 if the shape is not supported
 *                                  (needs D == 4, Wp == 8, M*128 B <= 192 KB: one HALF -- feature or gradient
 *                                  columns -- of 8 branches' codebooks per pass, stored so that the eight lanes of a
 *                                  shared-memory load phase hit eight different bank groups: conflict-free gathers).
 *   vqgnn_codes_group: codes_g[(k/G)*N + node][k%G] = codes[node, k]  (codes_g: [ceil(nb/G)][N][8] int16) for
 *                      the listed nodes (rows == NULL: all N) -- the group-major mirror of the code table.
 *   vqgnn_mp_fwd_tail: over a CSR holding ONLY tail entries (node = global node id):
 *       y[r]  += sum_e val[e] * feat_scale * O_k[code_k(node[e]), :D]      (y accumulates: run it AFTER
 *       gq[r]  = sum_e rval[e] * O_k[code_k(node[e]), D:2D]                 vqgnn_mp_fwd on the in-batch part)
 *       *info  = info_scale * sum_r <x[r], gq[r]>
 *     the tail sums are formed in a scratch copy with the same order-independent piece scheme as vqgnn_mp_fwd and
 *     added to y in one pass; ws: vqgnn_mp_fwd_tail_workspace_bytes(nnz, chunk, B, nb*D) bytes.
 *   d_nnz != NULL: the entry count lives on the device (vqgnn_plan_v1_build); nnz is then its upper bound. */
size_t vqgnn_mp_fwd_tail_workspace_bytes(int64_t nnz, int chunk, int64_t B, int C);
int vqgnn_mp_tail_group(int M, int D, int Wp);
/* Apply n (node, codes[nbc]) updates for branches [k0, k0 + nbc) to codes [N, nb] (and its group-major mirror
 * codes_g when not NULL); when a node is listed more than once the LAST entry wins.  Multi-GPU: the list is the
 * rank-ordered all-gather of every rank's re-assignments, so every replica ends up identical.
 * owner_ws: N int32 of scratch. */
int vqgnn_codes_apply_updates(const int32_t* nodes, const int16_t* new_codes, int64_t n, int nbc, int k0,
                              int16_t* codes, int nb, int64_t N, int16_t* codes_g, int G, int32_t* owner_ws,
                              void* stream);
int vqgnn_codes_group(const int16_t* codes, int nb, const int32_t* rows, int64_t n_rows, int64_t N, int G,
                      int16_t* codes_g, void* stream);
int vqgnn_mp_fwd_tail(const int32_t* rowptr, const int32_t* node, const float* val, const float* rval,
                      const int32_t* chunk_row, int chunk, int64_t nnz, const int32_t* d_nnz, int64_t B,
                      const float* x, int64_t ldx, const int16_t* codes_g, int64_t N, const float* O, int nb, int M, int D, int Wp,
                      float feat_scale, float info_scale, float* y, int64_t ldy, float* gq, int64_t ldgq,
                      float* info, void* ws, void* stream);

/* Device-side, synchronisation-free construction of the v1 batch plan from the reference's batch tuple
 * (deg_inv, A_BN (r, c, v), A_BB (r, c, v) | None, A_NB_v | None, batch_idx) of
 * vq_gnn_v1/utils/dataloader.py:64-86 -- what `mapper` (:144-192) rebuilds per branch per layer.
 *   tail part     (A_BN entries whose column is not a batch node; all of them when bb_r == NULL):
 *                 t_rowptr [B+1], t_node / t_val / t_rval [nnz] (rv == NULL -> zeros), *t_count = number of tail
 *                 entries (stays on the device), t_chunk_row [ceil(nnz / tail_chunk)]
 *   in-batch part (A_BB, + transpose if `symmetric` (GCN, to_symmetric), + self loops deg_inv if `self_loops`,
 *                 doubled when symmetric): nin = nbb * (1 + symmetric) + self_loops * B entries;
 *                 forward CSR i_rowptr / i_col / i_val / i_chunk_row and transposed b_rowptr / b_row / b_val /
 *                 b_chunk_row (chunk rows: ceil(nin / chunk) each)
 * A_BN must be row-sorted (the reference's loader emits CSR order): the tail part is a STABLE compaction of it, and
 * every in-batch CSR row is sorted by column, so the plan -- and with it the order of every fp32 sum in the
 * message-passing kernels -- does not depend on atomic ordering.
 * ws: vqgnn_plan_v1_workspace_bytes(N, B, nnz, nin) bytes. */
size_t vqgnn_plan_v1_workspace_bytes(int64_t N, int64_t B, int64_t nnz, int64_t nin);
int vqgnn_plan_v1_build(const int64_t* r, const int64_t* c, const float* v, const float* rv, int64_t nnz,
                        const int64_t* bb_r, const int64_t* bb_c, const float* bb_v, int64_t nbb,
                        const int64_t* batch_idx, const float* deg_inv, int64_t B, int64_t N, int symmetric,
                        int self_loops, int tail_chunk, int chunk, int32_t* t_rowptr, int32_t* t_node,
                        float* t_val, float* t_rval, int32_t* t_count, int32_t* t_chunk_row, int32_t* i_rowptr,
                        int32_t* i_col, float* i_val, int32_t* i_chunk_row, int32_t* b_rowptr, int32_t* b_row,
                        float* b_val, int32_t* b_chunk_row, void* ws, void* stream);

/* Transposed CSR restricted to columns < B of a CSR over R rows (the backward structure of a v2 batch graph):
 * browptr [B+1], brow / bval sized nnz (upper bound; the first *count entries are valid, each column's entries
 * sorted by source row), *count = number of entries with col < B (stays on the device: read it back lazily).
 * ws: vqgnn_csr_transpose_workspace_bytes(B, nnz). */
size_t vqgnn_csr_transpose_workspace_bytes(int64_t B, int64_t nnz);
int vqgnn_csr_transpose_lt(const int32_t* rowptr, const int32_t* col, const float* val, int64_t R, int64_t nnz,
                           int64_t B, int32_t* browptr, int32_t* brow, float* bval, int32_t* count, void* ws,
                           void* stream);

/* ---------------------------------------------------------------------------------------------
 * Mini-batch graph construction on the device from the RESIDENT normalised graph (CSR: rowptr int64 [N+1],
 * col int32 [nnz], val fp32 [nnz]) and the batch's node ids (int64 [B]) -- SURVEY.md section 8 f1.  Only the node
 * ids cross PCIe; the reference builds the batch graph on the CPU and ships an int64 COO (vq_gnn_v2/utils/misc.py:57-75).
 * All calls of one batch share one workspace of vqgnn_khop_workspace_bytes(N, N) bytes (it holds the node -> row table).
 *
 * v2, `_k_hop_subgraph` (vq_gnn_v2/dataloader.py:98-148): subset = [batch nodes ; out-of-batch 1-hop neighbours B',
 * ascending node id]; train keeps every edge inside the subset (R = B + B' rows), eval only the batch rows (R = B).
 *   vqgnn_khop_mark : batch_idx32 [B] (the ids as int32), tail_node [capacity N; first *T valid], *T = B'
 *   vqgnn_khop_count: out_rowptr [R+1] of the relabelled CSR, *nnz = its entry count      (after the host read T)
 *   vqgnn_khop_fill : out_col / out_val [nnz]; column ids < B = batch rows, >= B = tail entry (col - B); inside a
 *                     row the graph's stored order is kept (stable compaction)               (after the host read nnz)
 * v1, `__collate__` (vq_gnn_v1/utils/dataloader.py:64-86): the reference's own batch tuple in its own format
 * (int64 COO, row-sorted), i.e. exactly what vqgnn_plan_v1_build consumes:
 *   vqgnn_collate_v1_count: off_bn / off_bb [B+1] (row offsets of A_BN / A_BB), counts[0] = nnz(A_BN),
 *                           counts[1] = nnz(A_BB) (with_bb == 0: A_BB is not built, counts[1] = 0)
 *   vqgnn_collate_v1_fill : A_BN (bn_r, bn_c, bn_v), nb_v = A_NB_v = deg[row node] * v * deg_inv[col] (NULL: skip),
 *                           A_BB (bb_r, bb_c local ids, bb_v; NULL: skip), deg_inv_out [B] = deg_inv[batch nodes]. */
size_t vqgnn_khop_workspace_bytes(int64_t N, int64_t rows);
int vqgnn_khop_mark(const int64_t* rowptr, const int32_t* col, const int64_t* node_idx, int64_t B, int64_t N,
                    int32_t* batch_idx32, int32_t* tail_node, int32_t* T, void* ws, void* stream);
int vqgnn_khop_count(const int64_t* rowptr, const int32_t* col, const int64_t* node_idx, const int32_t* tail_node,
                     int64_t B, int64_t R, int64_t N, int32_t* out_rowptr, int32_t* nnz, void* ws, void* stream);
int vqgnn_khop_fill(const int64_t* rowptr, const int32_t* col, const float* val, const int64_t* node_idx,
                    const int32_t* tail_node, int64_t B, int64_t R, int64_t N, const int32_t* out_rowptr,
                    int32_t* out_col, float* out_val, void* ws, void* stream);
int vqgnn_collate_v1_count(const int64_t* rowptr, const int32_t* col, const int64_t* node_idx, int64_t B, int64_t N,
                           int with_bb, int32_t* off_bn, int32_t* off_bb, int32_t* counts, void* ws, void* stream);
int vqgnn_collate_v1_fill(const int64_t* rowptr, const int32_t* col, const float* val, const float* gdeg,
                          const float* gdeg_inv, const int64_t* node_idx, int64_t B, int64_t N, const int32_t* off_bn,
                          const int32_t* off_bb, int64_t* bn_r, int64_t* bn_c, float* bn_v, float* nb_v, int64_t* bb_r,
                          int64_t* bb_c, float* bb_v, float* deg_inv_out, void* ws, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Message passing, GAT, v2 ("B+B'") formulation: OurGATConv.forward/message (vq_gnn_v2/convs.py:165-266)
 * with vq_softmax == un-normalised exp (vq_gnn_v2/utils/vq_softmax.py:41-57), fused with the codeword
 * gather, the ones-column denominator and the batch-row normalisation of vq_gnn_v2/models.py:161-198.
 * Nodes of the batch graph: n < B batch rows (features x[n]), n >= B out-of-batch nodes (features = their
 * codewords O_k[code, :D]); every node carries an implicit extra column of ones (models.py:176-177).
 * att_l / att_r: [C+1] floats (the reference's [1,1,C+1] parameters).
 * ------------------------------------------------------------------------------------------- */

/* a_l[n] = <Xin[n], att_l>, a_r[n] = <Xin[n], att_r> for n < R (convs.py:188-190);
 * stat[0] = max_n a_l[n], stat[1] = max_n a_r[n] ("Trick 1" scale, convs.py:209-211). */
int vqgnn_gat_scores(int64_t R, int64_t B, const float* x, int64_t ldx, const int32_t* tail_node,
                     const int16_t* codes, const float* O, int nb, int M, int D, int Wp, const float* tail_feat,
                     int64_t ld_tail, const float* att_l, const float* att_r, float* a_l, float* a_r,
                     float* stat, void* stream);

/* sigma = sqrt(stat[0]^2+1) sqrt(stat[1]^2+1); w_ij = val * exp(leaky_relu((a_l[j]+a_r[i])/sigma, slope));
 * rows i < B:  den[i] = sum_j w_ij ; y[i, :C] = sum_j w_ij Xin[j, :C] / (den[i] + 1e-16)
 * rows i >= B: info += <sum_j w_ij Xin[j, :C], O_k[code(node(i-B)), D:2D]>   (un-normalised, models.py:198)
 * *info = info_scale * info.  Partition arguments as in vqgnn_mp_fwd.  tail_feat / tail_grad (here and in the
 * other two entry points): optional dense rows from vqgnn_tail_materialize, NULL = gather per use. */
int vqgnn_gat_fwd(const int32_t* rowptr, const int32_t* col, const float* val, const int32_t* chunk_row,
                  int chunk, int64_t nnz, int64_t R, int64_t B, const float* x, int64_t ldx,
                  const int32_t* tail_node, const int16_t* codes, const float* O, int nb, int M, int D, int Wp,
                  const float* tail_feat, int64_t ld_tail, const float* a_l, const float* a_r, const float* stat,
                  float negative_slope, float info_scale, float* y, int64_t ldy, float* den, float* info,
                  void* ws, void* stream);

/* The same forward for a batch graph whose codeword rows were materialised (vqgnn_tail_materialize: tail_feat, and
 * tail_grad when info is wanted), through the lean row-gather kernel of vqgnn_mp_fwd_rows with the GAT weights computed
 * per entry (csrc/mp_rows.cuh).  a_l / a_r / stat from vqgnn_gat_scores over all B + T nodes.  Needs C >= 16, C % 4 == 0,
 * chunk <= 256.  ws: 512 + 8 * ceil(ceil(nnz / chunk) * ceil(C / 128) / 4) bytes when info != NULL. */
int vqgnn_gat_fwd_rows(const int32_t* rowptr, const int32_t* col, const float* val, const int32_t* chunk_row,
                       int chunk, int64_t nnz, int64_t R, int64_t B, const float* x, int64_t ldx,
                       const float* tail_feat, int64_t T, const float* tail_grad, int64_t ld_tail, int C,
                       const float* a_l, const float* a_r, const float* stat, float negative_slope,
                       float info_scale, float* y, int64_t ldy, float* den, float* info, void* ws, size_t ws_bytes,
                       void* stream);


/* Backward of vqgnn_gat_scores + vqgnn_gat_fwd given dout = d loss / d y [B, C] and dinfo (device scalar or
 * NULL = 1); `out` is the forward's y.  Outputs:
 *   dyn [B, C]  = dout / (den + 1e-16): the gradient w.r.t. the un-normalised conv output, which is what the
 *                 reference's VQ hook receives (vq_gnn_v2/models.py:181-185);
 *   dx  [B, C]  (may be NULL), datt_l / datt_r [C+1];
 * scratch: dden [B], ds_l / ds_r [R].  Includes the gradient through the max-based scale (convs.py:209, in
 * the autograd graph) and of info_backward (tail rows weigh tail_scale * dinfo * gradient codeword). */
int vqgnn_gat_bwd(const int32_t* rowptr, const int32_t* col, const float* val, const int32_t* chunk_row,
                  int64_t nnz, int64_t R, const int32_t* browptr, const int32_t* brow, const float* bval,
                  const int32_t* bchunk_row, int64_t bnnz, int chunk, int64_t B, const float* x, int64_t ldx,
                  const int32_t* tail_node, const int16_t* codes, const float* O, int nb, int M, int D, int Wp,
                  const float* tail_feat, const float* tail_grad, int64_t ld_tail,
                  const float* att_l, const float* att_r, const float* a_l, const float* a_r, const float* stat,
                  float negative_slope, const float* out, int64_t ldo, const float* den, const float* dout,
                  int64_t lddo, float tail_scale, const float* dinfo, float* dyn, int64_t lddyn, float* dden,
                  float* ds_l, float* ds_r, float* dx, int64_t lddx, float* datt_l, float* datt_r,
                  void* stream);

/* ---------------------------------------------------------------------------------------------
 * Message passing, GAT, v1 ("B+M") formulation: vq_gnn_v1/models.py:143-233 + `mapper`
 * (vq_gnn_v1/utils/dataloader.py:144-192) + OurGATConv (vq_gnn_v1/convs.py).  Every branch is an independent
 * GAT over B batch rows and M codewords with D+1 = 5 columns and its own att_l / att_r ([nb, 5] stacked);
 * needs num_D = 4 and the add_flag codebook layout (Wp = 12: 4 feature | 4+1 gradient | pad).
 * CSR = the v1 plan (col < B in-batch incl. the self loop, col >= B tail entry with node id col - B; rval =
 * A_NB_v reverse values, may be NULL in eval).
 * ------------------------------------------------------------------------------------------- */

/* a_l / a_r [B, nb] (batch rows), cs [nb, M, 2] (codewords: c_l, c_r; features wu * O_k[m, :4]),
 * stat [nb, 2] = per-branch maxima of the l / r scores over batch rows and codewords (convs.py:209). */
int vqgnn_gat1_scores(int64_t B, const float* x, int64_t ldx, const float* O, int nb, int M, int D, int Wp,
                      float wu, const float* att_l, const float* att_r, float* a_l, float* a_r, float* cs,
                      float* stat, void* stream);
/* y [B, nb*4] = per-branch normalised outputs (models.py:209-210), den [B, nb] the ones-column sums,
 * *info = wu * sum over branches of <X_out_M, gradient codewords> (models.py:223) when info != NULL. */
int vqgnn_gat1_fwd(const int32_t* rowptr, const int32_t* col, const float* val, const float* rval,
                   const int32_t* chunk_row, int chunk, int64_t nnz, int64_t B, const float* x, int64_t ldx,
                   const int16_t* codes, const float* O, int nb, int M, int D, int Wp, float wu, const float* a_l,
                   const float* a_r, const float* cs, const float* stat, float negative_slope, float* y,
                   int64_t ldy, float* den, float* info, void* ws, void* stream);
/* Backward.  gy [B, nb*5] = d loss / d (un-normalised X_out_B) per branch, i.e. the VQ hook's gradient
 * (models.py:199-203, add_flag width D+1); dx [B, nb*4] (may be NULL); datt_l / datt_r [nb, 5];
 * scratch ds [B, nb, 2], dcs [nb, M, 2]. */
int vqgnn_gat1_bwd(const int32_t* rowptr, const int32_t* col, const float* val, const float* rval,
                   const int32_t* chunk_row, int chunk, int64_t nnz, int64_t B, const float* x, int64_t ldx,
                   const int16_t* codes, const float* O, int nb, int M, int D, int Wp, float wu,
                   const float* att_l, const float* att_r, const float* a_l, const float* a_r, const float* cs,
                   const float* stat, float negative_slope, const float* out, int64_t ldo, const float* den,
                   const float* dout, int64_t lddo, const float* dinfo, float* gy, float* ds, float* dcs,
                   float* dx, int64_t lddx, float* datt_l, float* datt_r, void* stream);

/* ---------------------------------------------------------------------------------------------
 * helpers
 * ------------------------------------------------------------------------------------------- */
int vqgnn_fill_zero(void* ptr, size_t bytes, void* stream);
/* codes[N, nb] <-> nb separate c_indices[N] tables (the reference layout) */
int vqgnn_codes_pack(const int16_t* const* h_tables, int nb, int64_t N, int16_t* codes, void* stream);
/* L2 flush helper for benchmarks: writes `bytes` of zeros to buf */
int vqgnn_flush_l2(void* buf, size_t bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VQGNN_H_ */
