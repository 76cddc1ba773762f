/* anemoi_b200.h — C ABI of libanemoi_b200.so: hand-written sm_100a kernels for the anemoi-models graph
 * message-passing forward (GraphConv / GraphTransformer processor blocks + grid<->mesh mappers).
 *
 * The reference (ecmwf/anemoi-core) has NO native code; every entry point below replaces a chain of
 * PyTorch / torch_geometric / Triton calls of the reference hot path.  Paths are relative to
 * /root/reference/models/src/anemoi/models/ .  INTEGRATION.md shows the ctypes binding and where the
 * reference's modules would call each symbol.
 *
 * Conventions
 *  - All pointers are DEVICE pointers owned by the caller (PyTorch) unless an entry point says HOST.  The library never allocates
 *    or frees device memory and keeps no global device state (host-side it only caches TMA descriptors) - with ONE explicit
 *    exception: the symmetric peer buffers of the multi-GPU exchange, created and destroyed by anemoi_b200_ipc_alloc / _free.
 *  - Every call only ENQUEUES work on `stream` (a cudaStream_t / CUstream passed as void*): no device
 *    synchronisation, CUDA-graph capturable, re-entrant, one caller thread per device.
 *  - Matrices are row-major with an explicit leading dimension in ELEMENTS (`ld*`), so column slices of wider
 *    buffers (q|k|v|self, x|aggregate) are passed without copies.
 *  - dtype codes: ANEMOI_F32 = 0, ANEMOI_BF16 = 1.  Biases, LayerNorm affine parameters, edge attributes and
 *    the lin_edge weights are always fp32.  Accumulation is always fp32.
 *  - Index arrays produced by anemoi_b200_csr_build are int32 (`src32`, `colptr32`); the reference-facing
 *    `colptr64` keeps the reference's int64 dtype (triton/utils.py:61-70) for bit-exact comparison.
 *  - Return value: 0 on success, negative on error (-1 bad argument, -2 CUDA error, -3 unsupported shape);
 *    anemoi_b200_last_error() returns a thread-local message.  Nothing throws or exits across the ABI.
 */
#ifndef ANEMOI_B200_H
#define ANEMOI_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ANEMOI_F32 0
#define ANEMOI_BF16 1

#if defined(__GNUC__)
#define ANEMOI_API __attribute__((visibility("default")))
#else
#define ANEMOI_API
#endif

/* anemoi_b200_linear flags */
#define ANEMOI_EPI_GELU 1 /* exact (erf) GELU after bias/gather-add, before residual  (torch.nn.GELU default) */
#define ANEMOI_EPI_G1_BF16 8  /* gather table g1 holds bf16 rows (ldg in bf16 elements) instead of fp32 */
#define ANEMOI_EPI_G2_BF16 16 /* same for g2: halves the L2 traffic of the gather-add epilogue (GraphConv's x_j = x_src[src] term) */
#define ANEMOI_EPI_NOPDL 4 /* launch without the programmatic-dependent-launch attribute (a launch that follows a cross-stream event wait) */
#define ANEMOI_EPI_REVERSE 2 /* scheduling hint, results unchanged: walk the row blocks from the last to the first, so that a GEMM
                              * consuming what the previous kernel wrote last starts where L2 still holds it (bf16 tcgen05 path) */

ANEMOI_API int anemoi_b200_abi_version(void);
ANEMOI_API const char* anemoi_b200_last_error(void);

/* -- integer path ------------------------------------------------------------------------------------------
 * Replaces triton/utils.py:25-70 `edge_index_to_csc` (colptr = index2ptr(dst); row = src) and the sortedness
 * check of distributed/khop_edges.py:43-48 for an edge_index [2,E] int64 that is already dst-sorted
 * (layers/graph_provider.py:185).  Built once per static graph instead of per layer per forward
 * (layers/block.py:779-782).
 *   edge_index : int64 [2, n_edges] row-major (row 0 = src, row 1 = dst)
 *   colptr64   : int64 [n_dst+1]  (may be NULL)      colptr32 : int32 [n_dst+1]
 *   src32      : int32 [n_edges]                     dst32 : int32 [n_edges] (may be NULL)
 *   status     : int32 [1], bit0 set if dst not non-decreasing, bit1 if an index is out of range
 */
ANEMOI_API int anemoi_b200_csr_build(const int64_t* edge_index, int64_t n_edges, int64_t n_src, int64_t n_dst, int64_t* colptr64,
                          int32_t* colptr32, int32_t* src32, int32_t* dst32, int32_t* status, void* stream);

/* -- LayerNorm ---------------------------------------------------------------------------------------------
 * torch.nn.LayerNorm / AutocastLayerNorm (layers/normalization.py:19-31) over the last `C` elements, eps as
 * given, optional affine (gamma/beta may be NULL), optional residual added AFTER the normalisation
 * (GraphConv: e' = LN(h) + e, conv.py:75-76; node_mlp(...) + x, block.py:393).
 * x is addressed as x[m*ldx + g*C + c] for m < M, g < groups (groups > 1: per-head q/k norm, block.py:655-660).
 */
ANEMOI_API int anemoi_b200_layer_norm(const void* x, int64_t ldx, int x_dtype, const float* gamma, const float* beta, const void* residual,
                           int64_t ldr, int r_dtype, void* y, int64_t ldy, int y_dtype, int64_t M, int64_t groups, int64_t C,
                           float eps, void* stream);

/* ConditionalLayerNorm (layers/normalization.py:34-94): y = LN(x) * (1 + cond . w_scale^T + b_scale) + (cond . w_bias^T + b_bias), the
 * LayerNorm itself without affine.  cond : fp32 [M, ldc >= Dc] (Dc <= 32), w_scale / w_bias : fp32 [C, Dc] (the `scale` / `bias` Linear
 * weights), b_scale / b_bias : fp32 [C].  The per-row scale and bias ([M, C] each in the reference) are formed inside the kernel. */
ANEMOI_API int anemoi_b200_cond_layer_norm(const void* x, int64_t ldx, int x_dtype, const float* cond, int64_t ldc, const float* w_scale,
                                const float* b_scale, const float* w_bias, const float* b_bias, void* y, int64_t ldy, int y_dtype, int64_t M,
                                int64_t C, int64_t Dc, float eps, void* stream);

/* -- Linear with fused epilogue ----------------------------------------------------------------------------
 * out[M,N] = epi( A[M,K] . W[N,K]^T ),  epi(acc) = [gelu]( acc + bias + g1[idx1[m]] + g2[idx2[m]] ) + residual
 * Replaces torch.nn.Linear (+ GELU, + residual add, + the torch.cat([x_i, x_j, e]) of conv.py:74 through the
 * gather-add terms: W.[x_i;x_j;e] = (W_i x_dst)[dst] + (W_j x_src)[src] + W_e e ).
 *   a_dtype == ANEMOI_BF16 : A, W bf16 -> tcgen05 (UMMA 128xBNx16, TMA SW128 operands, TMEM accumulators).
 *                            Requires lda % 8 == 0, ldw % 8 == 0, 16-byte aligned A and W.
 *   a_dtype == ANEMOI_F32  : A, W fp32 -> fp32 FFMA path (parity mode; no tensor cores, no TF32 rounding).
 *   bias  : fp32 [N] or NULL.  g1/g2 : fp32 (or, with ANEMOI_EPI_G1_BF16 / _G2_BF16, bf16) [*, ldg] gathered by int32 row indices idx1/idx2 [M], or NULL.
 *   residual : [M, ldr] of r_dtype or NULL.   out : [M, ldo] of o_dtype.
 *   ln_stats / ln_colsum (both or neither): LayerNorm of the A operand folded into the GEMM.  With W' = W * gamma (per input column),
 *     LN(x) W^T + b = rstd_m (x W'^T - mean_m colsum_n) + (b + W beta),  colsum_n = sum_k W'[n,k]:  the caller passes the RAW x as A,
 *     W' as W, (b + W beta) as bias, per-row (mean, rstd) from anemoi_b200_row_stats as ln_stats [M,2] and colsum [N]; the separate
 *     LayerNorm pass (one read + one write of the activations) disappears.
 *   stats_out (may be NULL): fp32 [M, ceil(N/64), 2]; receives per row and 64-column block (sum, sum of squares) of the output AS STORED.
 *     The GEMM that consumes `out` through a folded LayerNorm then passes this buffer as ln_stats with ln_parts = ceil(K/64),
 *     ln_dim = K (normalised width), ln_eps; (mean, rstd) are formed in its epilogue and the anemoi_b200_row_stats pass disappears too.
 *     ln_parts == 0: ln_stats is the [M,2] (mean, rstd) form.  The bf16 tcgen05 epilogue writes stats_out itself (plain / residual form);
 *     other paths run one extra pass over `out`.
 */
ANEMOI_API int anemoi_b200_linear(const void* A, int64_t lda, const void* W, int64_t ldw, int a_dtype, const float* bias, const float* g1,
                       const int32_t* idx1, const float* g2, const int32_t* idx2, int64_t ldg, const void* residual, int64_t ldr,
                       int r_dtype, void* out, int64_t ldo, int o_dtype, int64_t M, int64_t N, int64_t K, int flags, const float* ln_stats,
                       const float* ln_colsum, int64_t ln_parts, int64_t ln_dim, float ln_eps, float* stats_out, void* stream);

/* per-row LayerNorm statistics (mean, 1/sqrt(var + eps)), two-pass in registers: stats[m] = (mean, rstd), x [M, ldx] of x_dtype.
 * flags: ANEMOI_EPI_REVERSE = read the rows from the last to the first (scheduling hint, see anemoi_b200_linear; results unchanged). */
ANEMOI_API int anemoi_b200_row_stats(const void* x, int64_t ldx, int x_dtype, float* stats, int64_t M, int64_t C, float eps, int flags, void* stream);

/* -- GraphTransformer edge-softmax attention (forward) -------------------------------------------------------
 * Replaces layers/conv.py:103-147 (GraphTransformerConv, PyG softmax) and triton/gt.py:81-179 / :390-428
 * (`anemoi::graph_transformer_attention`), optionally fused with lin_edge (block.py:623-635):
 *   out[d,h,:] = sum_e alpha[e,h] (v[src_e,h,:] + eproj[e,h,:]) (+ add[d,h,:]),
 *   alpha = softmax over edges into d of  q[d,h,:].(k[src_e,h,:] + eproj[e,h,:]) / sqrt(Ch)
 * Zero in-degree rows give 0 (+ add), like gt.py:112-119.  fp32 online softmax.  Three forms of the edge term:
 *   (1) materialised: `e` = eproj [E, H*Ch] of `dtype` (the reference operator boundary);
 *   (2) in-kernel projection: edge_attr fp32 [E, lde] + w_edge fp32 [H*Ch, ldw_e] + b_edge -> eproj = w_edge.a + b_edge;
 *   (3) folded (the fast path, eproj is linear in a):  q.(W a + b) = (W^T q).a + const  and
 *       sum alpha (W a + b) = W (sum alpha a) + b, so the caller passes qw[d,h,:dp] = W_h^T q[d,h,:] (extra columns of the q GEMM)
 *       and receives abar[d,h,:dp] = sum_e alpha a_e, which it feeds to the projection GEMM through W; the kernel adds b_edge.
 *       The [E, H*Ch] edge tensor (the largest tensor of the layer in the reference) never exists.  Needs edge_dim <= 16
 *       and edge_attr rows zero-padded to lde >= 16.
 *   q [n_dst, ldq], k,v [n_src, ldk/ldv], add/out [n_dst, ld*], qw/abar [n_dst, >= heads*dp] of `dtype`; src32/colptr32 from csr_build.
 *   lse (nullable): fp32 [n_dst, heads] receives log sum_e exp(score_e) (0 for rows without edges, like the `m` output of
 *   triton/gt.py:112-119, 170-178); forms (1) / (2) only - it is what anemoi_b200_gt_attention_bwd consumes.
 *   flags: ANEMOI_EPI_REVERSE = the warps take their destination ranges from the last rows to the first (scheduling hint: the rows the
 *   producing GEMM wrote last are still in L2; results unchanged).
 */
ANEMOI_API int anemoi_b200_gt_attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* e,
                                 int64_t lde_proj, const float* edge_attr, int64_t lde, int64_t edge_dim, const float* w_edge,
                                 int64_t ldw_e, const float* b_edge, const void* qw, int64_t ldqw, void* abar, int64_t ldabar, int64_t dp,
                                 const int32_t* src32, const int32_t* colptr32, const void* add, int64_t ldadd, void* out, int64_t ldo,
                                 float* lse, int64_t n_dst, int64_t heads, int64_t ch, int dtype, int flags, void* stream);

/* -- GraphTransformer attention on destination tiles (folded form (3), bf16, Ch in {32, 64}) ------------------------
 * Same result as anemoi_b200_gt_attention_fwd form (3) (replaces layers/conv.py:103-147, triton/gt.py:81-179 and the lin_edge GEMM of
 * block.py:623-635), computed per tile of <= 16 consecutive destination rows and the <= 64 distinct source rows they name:
 * S = Q K^T and O = P V on the warp-level tensor cores (mma.sync m16n8k16) with every k / v row gathered ONCE per tile, the sparsity
 * mask and the edge term as an additive bias tile.  Meant for node orders with locality (layers/_reorder.py).
 *
 * anemoi_b200_attn_tile_plan is the HOST-side analysis step (HOST pointers, run once per static graph): from the dst-sorted CSR it
 * builds tile_meta [n_tiles, 4] int32 = {first dst row, first slot, first edge, rows | slots << 8 | edges << 16}, slot_src (source row
 * of every slot) and emeta [E] uint16 (slot of every edge inside its tile | row inside its tile << 8).  Capacities: tile_meta 4 * n_dst,
 * slot_src E, emeta E.  max_edges bounds the edges of a tile: <= 224 and <= 10752 / (4 dp) (the kernel stages dp fp32 attributes per edge
 * in shared memory).  Returns -3 when a destination has more than 64 distinct sources or max_edges edges (no plan; use
 * anemoi_b200_gt_attention_fwd).
 * The plan arrays are then copied to the device by the caller (tile_meta 16-byte aligned) and passed to the kernel entry point.
 */
ANEMOI_API int anemoi_b200_attn_tile_plan(const int32_t* src32_host, const int32_t* colptr32_host, int64_t n_src, int64_t n_dst,
                               int64_t max_edges, int32_t* tile_meta_host, int32_t* slot_src_host, uint16_t* emeta_host, int64_t* n_tiles,
                               int64_t* n_slots);
ANEMOI_API int anemoi_b200_gt_attention_tiled_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                       const float* edge_attr, int64_t lde, const float* b_edge, const void* qw, int64_t ldqw, void* abar,
                                       int64_t ldabar, int64_t dp, const int32_t* colptr32, const int32_t* tile_meta,
                                       const int32_t* slot_src, const uint16_t* emeta, int64_t n_tiles, const void* add, int64_t ldadd,
                                       void* out, int64_t ldo, int64_t n_dst, int64_t heads, int64_t ch, void* stream);

/* -- GraphConv tail: LayerNorm + residual + dst-segmented sum ------------------------------------------------
 * Replaces the tail of layers/conv.py:73-81: e'[i] = LN(h[i]) + e[i]  (written to e_new) and
 * out[d] = sum_{edges i into d} e'[i]  (scatter-sum; deterministic segmented reduction over the dst-sorted
 * edges, fp32 accumulation, no atomics).  h, e, e_new : [E, ld*]; out : [n_dst, ldo].
 */
ANEMOI_API int anemoi_b200_graphconv_ln_aggregate(const void* h, int64_t ldh, const float* gamma, const float* beta, const void* e, int64_t lde,
                                       void* e_new, int64_t ldn, const int32_t* colptr32, void* out, int64_t ldo, int64_t n_dst,
                                       int64_t C, float eps, int dtype, void* stream);

/* -- GraphConv, the whole operator as ONE kernel (north star: fused edge-gather -> edge-MLP -> segmented scatter-reduce) ------------
 * Replaces layers/conv.py:66-81 end to end for C in {16, 32, 64}, the widths at which the operator is HBM-bound:
 *   e_new[i] = LayerNorm(edge_mlp([x_dst[dst_i] ; x_src[src_i] ; e[i]])) + e[i],   out[d] = sum of e_new over the edges into d,
 * edge_mlp = Linear(3C, C), GELU, (Linear(C, C), GELU) x (n_layers - 2), Linear(C, C)   (layers/mlp.py:96-170 with the plain "mlp"
 * implementation and the exact-erf GELU).  Per tile of 128 consecutive dst-sorted edges the operand rows are gathered into shared
 * memory (cp.async, one tile ahead), the layers run on mma.sync (bf16; hidden activations stay in registers) or FFMA (fp32 parity mode)
 * against weights resident in shared memory, LayerNorm + residual are applied in place and the destination sums are formed from the tile
 * (fp32, edge order, no atomics).  No intermediate tensor touches HBM.
 *   x_src [n_src, lds], x_dst [n_dst, ldd], e / e_new [n_edges, lde / ldn], out [n_dst, ldo] of `dtype`, 16-byte aligned rows;
 *   weights : `dtype`, packed [C, 3C] row-major followed by (n_layers - 1) x [C, C];  biases : fp32 [n_layers, C] (zeros where a layer
 *   has none);  gamma / beta : fp32 [C] or NULL;  src32 / dst32 / colptr32 from anemoi_b200_csr_build.
 * Returns -3 (use anemoi_b200_linear + anemoi_b200_graphconv_ln_aggregate) for other widths, layer counts or alignments.
 */
ANEMOI_API int anemoi_b200_graphconv_fused(const void* x_src, int64_t lds, const void* x_dst, int64_t ldd, const void* e, int64_t lde,
                                const void* weights, const float* biases, int64_t n_layers, const float* gamma, const float* beta, void* e_new,
                                int64_t ldn, const int32_t* src32, const int32_t* dst32, const int32_t* colptr32, void* out, int64_t ldo,
                                int64_t n_dst, int64_t n_edges, int64_t C, float eps, int dtype, void* stream);

/* -- backward of the fused ops (training) -------------------------------------------------------------------------
 * gt_attention_bwd replaces triton/gt.py:182-376 / :451-556 for the materialised-e form (1) of anemoi_b200_gt_attention_fwd: given the
 * forward's out [n_dst, ldo] (the attention result, without `add`) and lse [n_dst, heads] (natural-log softmax normaliser, written by the
 * forward when its `lse` argument is non-null) and the cotangent dout, it returns dq, dk, dv and (when e is given) de; two deterministic
 * passes (dst-major over src32 / colptr32, src-major over the reverse CSR rev_ptr32 [n_src + 1] / rev_eid32 [E] = edge ids stably sorted by
 * source, with dst32 [E]); alpha_scratch / ds_scratch: fp32 [E, heads] workspaces.  Channels per head <= 256.
 * layer_norm_bwd: dx (and dres = the cotangent itself, nullable) of y = LayerNorm(x) * gamma + beta over `groups` groups of C <= 1024
 * channels per row, for the cotangent g[r] = dy[r] (nullable) + dz[idx[r]] (nullable; idx nullable = identity); per-block partial sums of
 * dgamma / dbeta go to partial [n_partial, 2, C] fp32 (the caller sums over blocks).  With dz = d out and idx = dst32 this is the backward
 * of anemoi_b200_graphconv_ln_aggregate (layers/conv.py:73-81).
 * gelu: mode 0  y = gelu(x);  mode 1  y = dy * gelu'(x)  (exact erf GELU);  mode 2  as mode 1 with Phi and phi through two exp2 (|error of
 *       gelu'| <= 1.3e-5; for bf16 cotangents, where the exact form is bound by instruction issue).
 */
ANEMOI_API int anemoi_b200_gt_attention_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const void* e,
                                 int64_t lde, const void* out, int64_t ldo, const void* dout, int64_t lddo, const float* lse,
                                 const int32_t* src32, const int32_t* colptr32, const int32_t* dst32, const int32_t* rev_ptr32,
                                 const int32_t* rev_eid32, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, void* de,
                                 int64_t ldde, float* alpha_scratch, float* ds_scratch, int64_t n_src, int64_t n_dst, int64_t heads, int64_t ch,
                                 int dtype, void* stream);
ANEMOI_API int anemoi_b200_layer_norm_bwd(const void* x, int64_t ldx, const float* gamma, const void* dy, int64_t lddy, const void* dz, int64_t lddz,
                               const int32_t* idx, void* dx, int64_t lddx, void* dres, int64_t lddr, float* partial, int64_t n_partial, int64_t M,
                               int64_t groups, int64_t C, float eps, int dtype, void* stream);
ANEMOI_API int anemoi_b200_gelu(const void* x, int64_t ldx, const void* dy, int64_t lddy, void* y, int64_t ldy, int64_t M, int64_t N, int mode,
                     int dtype, void* stream);

/* Backward of a row gather (training of GraphConv: the gradient of x_i = x_dst[dst], x_j = x_src[src]; PyTorch autograd's index_add_ in the
 * reference): out[n] = sum over j in [ptr32[n], ptr32[n+1]) of rows[eid32 ? eid32[j] : j].  With the dst-sorted edge list, ptr32 = colptr32 and
 * eid32 = NULL give the gradient of table[dst]; with the reverse CSR (edges grouped by source) the gradient of table[src].  Deterministic,
 * fp32 accumulation, no atomics.  rows : [*, ld], out : [n_out, ldo] of `dtype`, C a multiple of 16 bytes. */
ANEMOI_API int anemoi_b200_segment_sum(const void* rows, int64_t ld, const int32_t* ptr32, const int32_t* eid32, void* out, int64_t ldo, int64_t n_out,
                                       int64_t C, int dtype, void* stream);

/* Column sums (training: the bias gradient of a Linear, db = sum over rows of the cotangent; `dz.sum(0)` of PyTorch autograd in the reference):
 * out[c] = sum_r x[r, c], fp32, deterministic two-stage reduction without atomics.  x : [M, ld] of `dtype` (16-byte aligned rows), out : [N] fp32,
 * partial : [n_partial, round_up(N, 4)] fp32 scratch (16-byte aligned; the kernel uses up to min(n_partial, ...) row chunks; NULL / 0 = one chunk). */
ANEMOI_API int anemoi_b200_col_sum(const void* x, int64_t ld, float* out, float* partial, int64_t n_partial, int64_t M, int64_t N, int dtype,
                                   void* stream);

/* -- multi-GPU exchange over NVLink peer memory (one process per GPU, one NVSwitch box) ---------------------------------
 * Replaces, for the dst-range sharded forward, the NCCL collectives of the reference: `halo_exchange` (distributed/graph.py:466-484,
 * layers/block.py:1159-1169) and the source-row all-gather `sync_tensor` (graph.py:227-240).  Ranks write the rows their peers need
 * straight into the peers' tables through CUDA-IPC mapped pointers; all three steps are ordinary kernels on `stream`, so the exchange
 * is captured in the same CUDA graph as the compute around it.
 *   ipc_alloc   cudaMalloc + zero `bytes`, return the pointer and its 64-byte cudaIpcMemHandle_t (to be sent to the peers by the
 *               caller, e.g. torch.distributed.all_gather); ipc_open maps a peer's handle; ipc_close / ipc_free undo them.
 *   A CHANNEL is one 256-byte control block per rank in symmetric memory; ctl_ptrs [world] (HOST array) holds every rank's block as
 *   seen from this process (own rank: the local pointer).  Per exchange, in this order on every rank:
 *   peer_rendezvous  exchange number c = ++counter (on the device); tell every peer "started c", wait until every peer has: their
 *                    consumers of exchange c-1 are done, the tables may be overwritten.
 *   halo_push        rows[send_idx[j], :row_bytes] for j in [send_off[p], send_off[p+1]) go to dst_ptrs[p] + (j - send_off[p]) *
 *                    dst_ld_bytes in rank p (peer-mapped address, HOST array [world]); then "rows of c from me have arrived" to all peers.
 *                    send_idx / send_off are DEVICE int32 arrays (send_off has world + 1 entries).
 *   halo_wait        wait until the rows of exchange c from every peer have arrived.
 * A peer that does not show up within ~10 s traps the waiting kernel (error at the next synchronisation) instead of hanging.
 */
ANEMOI_API int anemoi_b200_ipc_alloc(int64_t bytes, void** dev_ptr, void* handle64);
ANEMOI_API int anemoi_b200_ipc_open(const void* handle64, void** dev_ptr);
ANEMOI_API int anemoi_b200_ipc_close(void* dev_ptr);
ANEMOI_API int anemoi_b200_ipc_free(void* dev_ptr);
ANEMOI_API int anemoi_b200_peer_rendezvous(const uint64_t* ctl_ptrs_host, int64_t world, int64_t rank, void* stream);
ANEMOI_API int anemoi_b200_halo_push(const void* rows, int64_t ld_bytes, int64_t row_bytes, const int32_t* send_idx, const int32_t* send_off,
                          int64_t n_send, const uint64_t* dst_ptrs_host, int64_t dst_ld_bytes, const uint64_t* ctl_ptrs_host, int64_t world,
                          int64_t rank, void* stream);
ANEMOI_API int anemoi_b200_halo_wait(const uint64_t* ctl_ptrs_host, int64_t world, int64_t rank, void* stream);

/* -- helpers -----------------------------------------------------------------------------------------------
 * cast/copy a [M, K] matrix into a [M, ldo] buffer (zero-filling columns K..Kpad-1), any of f32/bf16 <-> f32/bf16,
 * optionally gathering rows: out[m] = in[idx[m]] (idx int32, may be NULL).
 */
ANEMOI_API int anemoi_b200_cast_pad(const void* in, int64_t ldi, int i_dtype, const int32_t* idx, void* out, int64_t ldo, int o_dtype, int64_t M,
                         int64_t K, int64_t Kpad, void* stream);

/* out[m, c] = a[m, c] + b[m, c] — the latent skip connection around the processor
 * (models/encoder_processor_decoder.py:295-296), any mix of f32 / bf16, fp32 add. */
ANEMOI_API int anemoi_b200_add(const void* a, int64_t lda, int a_dtype, const void* b, int64_t ldb, int b_dtype, void* out, int64_t ldo,
                               int o_dtype, int64_t M, int64_t C, void* stream);

/* Gated feed-forward layer (layers/mlp.py:38-53 GatedMLPLayer; mlp_implementation = glu / swiglu / geglu / reglu):
 * out[m, j] = act(in[m, j]) * in[m, H + j] for j < H, where in = [gate_proj(x) | value_proj(x)] comes out of ONE anemoi_b200_linear call on
 * the row-concatenated weights.  act: 0 sigmoid, 1 silu, 2 exact-erf gelu, 3 relu.  in : [M, ldi >= 2H], out : [M, ldo >= H], same dtype. */
ANEMOI_API int anemoi_b200_glu_combine(const void* in, int64_t ldi, void* out, int64_t ldo, int64_t M, int64_t H, int act, int dtype, void* stream);

/* Backward of anemoi_b200_glu_combine (training with the gated MLP variants; PyTorch autograd of layers/mlp.py:38-53 in the reference):
 * din[m, j] = dy[m, j] * in[m, H + j] * act'(in[m, j]),  din[m, H + j] = dy[m, j] * act(in[m, j]).  in / din : [M, >= 2H], dy : [M, >= H]. */
ANEMOI_API int anemoi_b200_glu_combine_bwd(const void* in, int64_t ldi, const void* dy, int64_t lddy, void* din, int64_t ldd, int64_t M, int64_t H,
                                           int act, int dtype, void* stream);

/* fp32 Linear on the bf16 tensor cores (the fp32 parity mode's GEMMs; reference: torch.nn.Linear in fp32, layers/mlp.py:97-179):
 * x = x1 + x2 + x3 in bf16 parts (24 mantissa bits); products of bf16 numbers are exact in fp32, so sum_{i+j<=4} a_i w_j with fp32
 * accumulation is an fp32-grade product.  out[m, s*K + k] = part pa[s] (weight_side = 0: 3 2 1 2 1 1) or pw[s] (weight_side = 1:
 * 1 2 3 1 2 1) of in[m, k] - smallest products first, the tensor core's accumulation truncates; anemoi_b200_linear on the two [*, 6K] bf16 operands then returns A.W^T to ~1e-6 relative.
 * in : fp32 [M, ldi >= K], K % 4 == 0; out : bf16 [M, ldo >= 6K]. */
ANEMOI_API int anemoi_b200_split_bf16x3(const float* in, int64_t ldi, void* out, int64_t ldo, int64_t M, int64_t K, int weight_side, void* stream);

/* -- model glue either side of the path -------------------------------------------------
 * Replaces `_assemble_input` (models/encoder_processor_decoder.py:98-127): einops.rearrange(x, "batch time ensemble grid vars ->
 * (batch ensemble grid) (time vars)") + torch.cat with the node attributes (+ the autocast cast in front of the embedding Linear).
 * x : fp32 [B, T, E, G, V] contiguous; attrs : fp32 [attr_rows, A] (row r of the output reads attrs[r % attr_rows]), may be NULL when
 * A == 0; out : [B*E*G, ldo] of o_dtype, columns [0, T*V) = data, [T*V, T*V+A) = attributes, [T*V+A, Kpad) = 0.
 */
ANEMOI_API int anemoi_b200_assemble_input(const float* x, int64_t B, int64_t T, int64_t E, int64_t G, int64_t V, const float* attrs, int64_t A,
                               int64_t attr_rows, void* out, int64_t ldo, int64_t Kpad, int o_dtype, void* stream);

/* Replaces `_assemble_output` (models/encoder_processor_decoder.py:129-163) with the SkipConnection residual (layers/residual.py:60-81)
 * and the ReLU / LeakyReLU boundings (layers/bounding.py:81-94) fused:
 *   y[b, t, e, g, v] = bound_v(dec[(b e g), t*V_out + v] + (skip_src[v] >= 0 ? x[b, step, e, g, skip_src[v]] : 0))
 * dec : [B*E*G, ldd] of d_dtype; x : fp32 [B, T_in, E, G, V_in] (may be NULL: no residual); skip_src : int32 [V_out] input-variable index
 * of each output variable or -1 (may be NULL); bound : int32 [V_out] 0 = none, 1 = relu, 2 = leaky_relu(0.01) (may be NULL);
 * y : fp32 [B, T_out, E, G, V_out] contiguous.
 */
ANEMOI_API int anemoi_b200_assemble_output(const void* dec, int64_t ldd, int d_dtype, const float* x, int64_t B, int64_t T_in, int64_t E, int64_t G,
                                int64_t V_in, int64_t step, const int32_t* skip_src, const int32_t* bound, float* y, int64_t T_out,
                                int64_t V_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ANEMOI_B200_H */
