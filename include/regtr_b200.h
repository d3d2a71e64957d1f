/*
 * regtr_b200 -- C ABI of the B200-native RegTR correspondence-prediction hot path.
 *
 * The reference (yewzijian/RegTR) has no FFI layer: its boundary for this path is
 * the nn.Module surface (src/models/regtr.py:104-235) over ATen ops plus two
 * un-vendored CUDA libraries.  Each entry point below replaces one reference
 * operation (file:line cited per function; paths relative to /root/reference/src).
 * INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless marked "host";
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, no
 *     call synchronises, allocates or frees (workspaces are caller-owned and sized
 *     by the matching *_ws_bytes function);
 *   - stacked clouds are described by int32 prefix offsets `offs[n_clouds + 1]`
 *     (device memory) so that data-dependent level sizes never cross to the host
 *     inside the pyramid; `*_cap` arguments are host-known capacities (upper
 *     bounds of offs[n_clouds]) used only to size grids and buffers;
 *   - return value: 0 on success, REGTR_ERR_* (<0) on a rejected argument,
 *     -(1000 + cudaError_t) when a launch fails.  Data-dependent failures
 *     (coordinates outside the +-32767-cell key range) are reported through the
 *     device status word, see regtr_status_*.
 */
#ifndef REGTR_B200_H_
#define REGTR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define REGTR_OK 0
#define REGTR_ERR_ARG (-1)        /* null pointer / negative size / unsupported shape */
#define REGTR_ERR_WORKSPACE (-2)  /* workspace too small */
#define REGTR_ERR_UNSUPPORTED (-3)

#define REGTR_STATUS_KEY_RANGE 1u /* a voxel / cell coordinate left the 16-bit key range */
#define REGTR_STATUS_CAPACITY 2u  /* a capacity-bounded output (sub-sampled level) overflowed; results truncated */
#define REGTR_STATUS_GRID 4u      /* voxel bounding box beyond the dense-grid budget: redo with regtr_grid_subsample_sorted */

int regtr_version(void);                 /* ABI version, currently 1 */
const char* regtr_build_info(void);      /* host pointer: arch + compile flags string */

/* ---- pyramid pre-processing ------------------------------------------------------- */

/* Voxel-grid barycentre sub-sampling.
 * Replaces batch_grid_subsampling_kpconv_gpu (models/backbone_kpconv/kpconv.py:213-240,
 * i.e. MinkowskiEngine 0.5.4 SparseTensor(UNWEIGHTED_AVERAGE)) with the deterministic
 * rules of DESIGN.md: voxel = floor(p / dl) (IEEE fp32 division), output ordered by
 * ascending (cloud, vx, vy, vz), barycentre = fp32 sum in ascending input index / count.
 * xyz (n_cap,3) f32; offs (n_clouds+1) i32; out_xyz (out_cap,3); out_offs (n_clouds+1).
 * out_cap may be smaller than n_cap (static-shape pipelines): on overflow the output is
 * truncated memory-safely and REGTR_STATUS_CAPACITY is raised.
 * status: device uint32 word, OR-ed with REGTR_STATUS_* on data-dependent errors. */
size_t regtr_grid_subsample_ws_bytes(int n_cap, int n_clouds);
size_t regtr_grid_subsample_state_bytes(int n_cap);
int regtr_grid_subsample(const float* xyz, const int32_t* offs, int n_clouds, int n_cap, float dl,
                         float* out_xyz, int out_cap, int32_t* out_offs, uint32_t* status,
                         void* ws, size_t ws_bytes, void* state, size_t state_bytes, void* stream);
/* regtr_grid_subsample sorts by COUNTING over a dense voxel grid spanning each cloud's bounding box (hand-written
 * kernels only; budget: 16 cells per point of capacity, at least 2^18).  `state`: regtr_grid_subsample_state_bytes
 * bytes, ZERO before the first call and owned by this op between calls (every call leaves it zero).  A box beyond
 * the budget raises REGTR_STATUS_GRID; regtr_grid_subsample_sorted (stable library radix sort of (key, index)
 * pairs, any extent) gives the same result for such inputs. */
size_t regtr_grid_subsample_sorted_ws_bytes(int n_cap);
int regtr_grid_subsample_sorted(const float* xyz, const int32_t* offs, int n_clouds, int n_cap, float dl,
                                float* out_xyz, int out_cap, int32_t* out_offs, uint32_t* status,
                                void* ws, size_t ws_bytes, void* stream);

/* Uniform cell list over a stacked point set (search structure for regtr_ball_query).
 * `grid` is an opaque caller-owned buffer of regtr_cellgrid_bytes(n_cap) bytes; `order`
 * (n_cap) i32, optional, receives the cell-sorted permutation of the points (a spatially
 * coherent processing order for queries drawn from the same set). */
size_t regtr_cellgrid_bytes(int n_cap);
size_t regtr_cellgrid_ws_bytes(int n_cap);
size_t regtr_cellgrid_state_bytes(int n_cap);   /* ZERO before the first call; every call leaves it zero */
int regtr_cellgrid_build(const float* xyz, const int32_t* offs, int n_clouds, int n_cap, float cell,
                         void* grid, int32_t* order, uint32_t* status,
                         void* ws, size_t ws_bytes, void* state, size_t state_bytes, void* stream);

/* Fixed-radius neighbour search, first K supports in ascending index order.
 * Replaces batch_neighbors_kpconv_gpu (kpconv.py:261-288: pytorch3d 0.6.0 packed_to_padded +
 * ball_query + re-packing): per query keep support j (ascending) while
 * ((dx*dx + dy*dy) + dz*dz) < r*r in fp32 without FMA contraction; pad with the total
 * support count.  `s_grid` must have been built over (s, s_offs) with capacity s_cap and
 * cell >= radius.
 * q_order (optional, nq_cap): processing order of the queries (a permutation of [0,nq_cap)).
 * out_idx32 / out_idx64 (nq_cap,K): either may be NULL; rows of capacity padding
 * (>= q_offs[n_clouds]) are filled with the shadow index. */
int regtr_ball_query(const float* q, const int32_t* q_offs, const int32_t* q_order,
                     const float* s, const int32_t* s_offs, const void* s_grid,
                     int n_clouds, int nq_cap, int s_cap, int K, float radius,
                     int32_t* out_idx32, int64_t* out_idx64, void* stream);

/* ---- KPConv encoder --------------------------------------------------------------- */

/* Rigid KPConv, linear influence, 'sum' aggregation.
 * Replaces KPConv.forward (models/backbone_kpconv/kpconv_blocks.py:269-414):
 *   out[n] = (1/max(1,#{k: sum_c x[idx[n,k]] > 0})) * sum_p (sum_k h(n,k,p) x[idx[n,k]]) W[p]
 *   h = max(0, 1 - |s[idx[n,k]] - q[n] - kp[p]| / extent); idx == Ns is the shadow neighbour.
 * q (Nq,3) s (Ns,3) idx (Nq,K) i32, x (Ns,Cin) f32, W (P,Cin,Cout) f32, kp (P,3), out (Nq,Cout).
 * P must be 15.  Cin in {1..16} or 32/64/128/256.
 * nq_dev / ns_dev (optional, device int32): actual query / support counts when Nq / Ns are
 * capacities (static-shape pipelines); rows >= *nq_dev are padding (zeroed up to the next multiple of
 * 128, untouched beyond -- consumers work in 128-row tiles); the shadow index is *ns_dev.
 * ws: regtr_kpconv_fwd_ws_bytes(Nq, Ns, Cin, Cout) bytes (aggregated features + row flags + the split
 * transposed weights + the GEMM's workspace; the contraction runs on regtr_gemm_tf32x3 -- no library GEMM).
 * regtr_kpconv_ws_bytes(Nq, Ns, Cin) is the part regtr_kpconv_aggregate alone needs (wf + row flags). */
size_t regtr_kpconv_ws_bytes(int Nq, int Ns, int Cin);
size_t regtr_kpconv_fwd_ws_bytes(int Nq, int Ns, int Cin, int Cout);
int regtr_kpconv_fwd(const float* q, const float* s, const int32_t* idx, const float* x,
                     const float* W, const float* kp, int Nq, int Ns, const int32_t* nq_dev,
                     const int32_t* ns_dev, int K, int Cin, int Cout,
                     float extent, float* out, void* ws, size_t ws_bytes, void* stream);

/* Gather + influence + aggregation stage alone: wf (Nq, 15*Cin), already divided by the
 * neighbour count.  (The "neighbour gather" kernel of the north star.)  rowflag_ws (Ns bytes):
 * flags[r] = (sum_c x[r,c] > 0); computed here unless flags_ready != 0 (then it must already hold
 * them, e.g. from regtr_instnorm_act's rowflag_out). */
int regtr_kpconv_aggregate(const float* q, const float* s, const int32_t* idx, const float* x,
                           const float* kp, int Nq, int Ns, const int32_t* nq_dev, const int32_t* ns_dev,
                           int K, int Cin, float extent,
                           float* wf, uint8_t* rowflag_ws, int flags_ready, void* stream);

/* max over the K gathered rows with a zero shadow row.  Replaces max_pool
 * (kpconv_blocks.py:127-143).  x (Ns,C), idx (Nq,K) i32 -> out (Nq,C). */
int regtr_max_pool(const float* x, const int32_t* idx, int Nq, int Ns, const int32_t* ns_dev, int K, int C,
                   float* out, void* stream);

/* Per-cloud InstanceNorm1d(affine=False, eps) over the points of each cloud, optional
 * residual add, optional LeakyReLU.  Replaces BatchNormBlock.forward + nn.LeakyReLU
 * (kpconv_blocks.py:497-519, 546-561, 646, 741):  out = act(norm(x) + res).
 * x (n,C); offs (n_clouds+1) i32 device; n_cap >= offs[n_clouds]; res optional (n,C);
 * slope < 0 disables the activation.  In-place (out == x) is allowed.  Rows in [offs[n_clouds], n_cap)
 * are padding: zeroed up to the next multiple of 128, untouched beyond.
 * rowflag_out (optional, n_cap bytes, C/4 a power of two <= 32): flags[r] = (sum_c out[r,c] > 0),
 * the neighbour-count predicate of the KPConv that consumes `out`.
 * counters (optional): regtr_instnorm_counter_bytes(n_clouds, C) bytes of int32, ZERO before the first call
 * and owned by this op between calls (every call leaves them zero); with them the statistics are
 * finalised by the last statistics block (one launch fewer), without them by a separate kernel. */
size_t regtr_instnorm_ws_bytes(int n_cap, int n_clouds, int C);
size_t regtr_instnorm_counter_bytes(int n_clouds, int C);
int regtr_instnorm_act(const float* x, const int32_t* offs, int n_clouds, int n_cap, int C, float eps,
                       const float* res, float slope, float* out, uint8_t* rowflag_out,
                       void* ws, size_t ws_bytes, int32_t* counters, void* stream);

/* The apply pass alone: out = act((x - mean) * rstd + res) with stats (n_clouds, C, 2) = (mean, rstd) produced
 * elsewhere -- by regtr_gemm_tf32x3_instats, which accumulates them in the GEMM epilogue so that the Linear ->
 * InstanceNorm pairs of the KPConv blocks (kpconv_blocks.py:546-561, 401-406 + 497-519) never re-read their
 * output for the statistics.  rowflag_out as in regtr_instnorm_act. */
int regtr_instnorm_apply(const float* x, const int32_t* offs, int n_clouds, int n_cap, int C, const float* stats,
                         const float* res, float slope, float* out, uint8_t* rowflag_out, void* stream);

/* ---- dense layers ---------------------------------------------------------------- */

/* x = hi + lo with both halves exactly representable in TF32 (low 13 mantissa bits zero);
 * used to pre-split weight matrices for regtr_gemm_tf32x3. */
int regtr_split_tf32(const float* x, long long n, float* hi, float* lo, void* stream);

/* fp32-accurate GEMM on the tcgen05 tensor cores (3xTF32):
 *   C[M,N] = act(A[M,K] @ B[N,K]^T + bias[N] + R[M,N]),  B given pre-split as B_hi / B_lo.
 * Replaces nn.Linear (kpconv_blocks.py:546, regtr.py:36/145, transformers.py:95-101,
 * regtr.py:404-411) and the KPConv weight contraction (kpconv_blocks.py:401-406).
 * Row-major fp32; lda/ldb multiples of 4 and 16-byte aligned bases (TMA); bias / R optional;
 * m_dev (optional device int32): actual row count when M is a capacity; relu != 0 applies ReLU.
 * ws: regtr_gemm_ws_bytes(M,N,K) bytes (deterministic split-K planes for skinny long-K shapes;
 * also limits the tensor cores' truncating fp32 accumulation to short K runs). */
size_t regtr_gemm_ws_bytes(int M, int N, int K);
int regtr_gemm_tf32x3(const float* A, int lda, const float* B_hi, const float* B_lo, int ldb,
                      float* C, int ldc, const float* bias, const float* R, int ldr,
                      int M, int N, int K, const int32_t* m_dev, int relu,
                      void* ws, size_t ws_bytes, void* stream);

/* C = A @ B^T (no bias / residual / activation) PLUS the per-cloud InstanceNorm statistics of C:
 * stats (n_clouds, N, 2) = (mean, 1/sqrt(biased var + eps)) over the rows [offs[c], offs[c+1]) of each column,
 * without a second pass over C: every epilogue warp stores the column sums / sums of squares of its 32 rows
 * (fixed shuffle tree) to `part`, and a small second kernel adds the partials of each cloud in a fixed order
 * (fp64) -- no atomics, run-to-run bit-identical.  N % 32 == 0.
 * part: regtr_instnorm_part_bytes(M, N) bytes of scratch (8-byte aligned; contents irrelevant on entry). */
size_t regtr_instnorm_part_bytes(int M, int N);
int regtr_gemm_tf32x3_instats(const float* A, int lda, const float* B_hi, const float* B_lo, int ldb,
                              float* C, int ldc, int M, int N, int K, const int32_t* m_dev,
                              const int32_t* offs, int n_clouds, float eps, void* part, float* stats,
                              void* ws, size_t ws_bytes, void* stream);

/* ---- transformer ------------------------------------------------------------------ */

/* 3-D sine position embedding.  Replaces PositionEmbeddingCoordsSine.forward
 * (models/transformer/position_embedding.py:29-50).  dim_t (n_freq) f32 is the reference's
 * `temperature ** (2*(i//2)/n_freq)` table; out (n, d_model), zero padded. */
int regtr_pos_embed_sine(const float* xyz, int n, const float* dim_t, int n_freq, int d_model,
                         float scale, float* out, void* stream);

/* LayerNorm over the last dim with optional position add:  y = LN(x)*g + b ;
 * y_pos = y + pos.  Replaces nn.LayerNorm + with_pos_embed (transformers.py:117-119,
 * 194-196, 213-215, 232).  Any of y / y_pos may be NULL.  n_dev (optional, device): the real row count
 * when n is a capacity; rows beyond it are left untouched. */
int regtr_layernorm_pos(const float* x, const float* gamma, const float* beta, const float* pos,
                        int n, const int32_t* n_dev, int E, float eps, float* y, float* y_pos, void* stream);

/* Device-side attention problem table for a (src x B, tgt x B) token stack with cloud offsets
 * offs (2B+1): plan (6, 2B+1) i32 rows = q_start, q_len, cross k_start, cross k_len (the cross
 * partner of src_b is tgt_b and vice versa; regtr.py:156-166's key-padding masks made explicit), then the
 * exclusive prefix of the number of 64-query / 128-query tiles per problem (entry 2B = total): the `tile_base`
 * tables of the attention kernels below. */
int regtr_attention_plan(const int32_t* offs, int B, int32_t* plan, void* stream);

/* Variable-length multi-head attention core, fp32:  O = softmax(Q K^T * scale) V per head.
 * Replaces the attention core of nn.MultiheadAttention as called at
 * transformers.py:197-226 (key-padding masks become explicit (start,len) ranges).
 * Problem i attends queries rows [q_start[i], q_start[i]+q_len[i]) of Q to key rows
 * [k_start[i], k_start[i]+k_len[i]) of K/V.  Q/K/V/O are row-major with leading
 * dimensions ldq/ldk/ldv/ldo (floats); head h uses columns [h*head_dim, (h+1)*head_dim).
 * head_dim must be 32.  max_q_len: host upper bound of q_len[]. */
int regtr_mha_varlen_fwd(const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv,
                         float* O, int ldo, const int32_t* q_start, const int32_t* q_len,
                         const int32_t* k_start, const int32_t* k_len, int n_problems,
                         int max_q_len, const int32_t* tile_base, int max_tiles,
                         int n_heads, int head_dim, float scale, void* stream);
/* tile_base (optional, n_problems + 1, device): exclusive prefix of ceil(q_len / 64) with the total last; the launch
 * then covers max_tiles (a host bound of that total, e.g. capacity / 64 + n_problems) linear tiles instead of
 * ceil(max_q_len / 64) tiles per problem -- capacity-shaped launches know the per-problem lengths on the device only. */

/* CorrespondenceDecoder.simple_attention (regtr.py:316-351, the `direct_regress_coor: False` branch):
 * single-head attention whose values are the key coordinates,
 *   out[l, q] = sum_k softmax_k(Qp[l, q] . Kp[l, k] * scale) xyz[k]          (fp32)
 * for all n_layers decoder inputs at once.  Qp/Kp: (n_layers * n_rows, D) row-major with leading
 * dimension ld -- the q_proj / k_proj outputs; row l*n_rows + t belongs to token t of layer l.
 * xyz (n_rows, 3): token coordinates; out (n_layers * n_rows, 3).  Problem tables as for
 * regtr_mha_varlen_fwd (token ranges, shared by all layers).  D % 4 == 0. */
int regtr_corr_decode_fwd(const float* Qp, const float* Kp, int ld, const float* xyz, float* out,
                          const int32_t* q_start, const int32_t* q_len, const int32_t* k_start,
                          const int32_t* k_len, int n_problems, int max_q_len, int n_layers, int n_rows,
                          int D, float scale, void* stream);

/* Tensor-core attention core (bf16 operands, fp32 TMEM accumulation, fp32 softmax): the "fast"
 * precision mode of the same nn.MultiheadAttention core (transformers.py:197-226), head_dim 32.
 * Inputs are produced by regtr_gemm_tf32x3_qkv_bf16 (the packed in-projection with a bf16
 * epilogue): QK [n_tokens, 2E] bf16 row-major (q | k), Vt [E, ld_vt] bf16 (v transposed, columns
 * >= n_tokens zero).  O [n_tokens, E] fp32.  Problem tables as for regtr_mha_varlen_fwd. */
int regtr_gemm_tf32x3_qkv_bf16(const float* A, int lda, const float* B_hi, const float* B_lo, int ldb,
                               const float* bias, int M, int N, int K, int split, void* qk_out, int ld_qk,
                               void* vt_out, int ld_vt, const int32_t* m_dev, void* stream);
int regtr_mha_bf16_tc_fwd(const void* QK, int ld_qk, const void* Vt, int ld_vt, int n_tokens, float* O,
                          int ldo, const int32_t* q_start, const int32_t* q_len, const int32_t* k_start,
                          const int32_t* k_len, int n_problems, int max_q_len, int n_heads, int head_dim,
                          float scale, void* stream);

/* fp32-accurate attention core on the tcgen05 tensor cores (the parity mode of the same nn.MultiheadAttention core,
 * transformers.py:197-226): S = Q K^T and O = P V as 3xTF32 tcgen05.mma with TMA-fed operands, P kept in tensor
 * memory, fp32 softmax.  Inputs from regtr_gemm_tf32x3_qkv_split, the packed in-projection (N = 3E: q | k | v) whose
 * epilogue writes every value as its two TF32 halves: qk4 [n_tokens, 4E] fp32 = [Q_hi | Q_lo | K_hi | K_lo] with q
 * pre-multiplied by qscale (pass softmax_scale * log2(e)); vt2 [2E, ld_vt] fp32 = v transposed, hi rows then lo rows
 * (columns >= the real token count must be finite, e.g. zero).  O [n_tokens, E] fp32.  head_dim must be 32.
 * Problem tables as for regtr_mha_varlen_fwd. */
int regtr_gemm_tf32x3_qkv_split(const float* A, int lda, const float* B_hi, const float* B_lo, int ldb,
                                const float* bias, int M, int N, int K, int E, float qscale, float* qk4, int ld4,
                                float* vt2, int ld_vt, const int32_t* m_dev, void* stream);
int regtr_mha_tf32_tc_fwd(const float* qk4, int ld4, const float* vt2, int ld_vt, int n_tokens, float* O, int ldo,
                          const int32_t* q_start, const int32_t* q_len, const int32_t* k_start,
                          const int32_t* k_len, int n_problems, int max_q_len, const int32_t* tile_base,
                          int max_tiles, int n_heads, int head_dim, void* stream);
/* (tile_base / max_tiles as for regtr_mha_varlen_fwd, with 128-query tiles.) */

/* ---- pose ------------------------------------------------------------------------- */

/* Weighted Kabsch.  Replaces compute_rigid_transform (utils/se3_torch.py:108-154):
 * problem i uses rows [offs[i], offs[i+1]) of a,b (n,3) and w (n); T (n_problems,3,4),
 * T*a = b.  One warp per problem, fp64 accumulation, one-sided Jacobi 3x3 SVD. */
int regtr_kabsch_fwd(const float* a, const float* b, const float* w, const int32_t* offs,
                     int n_problems, float* T, void* stream);

/* Fused correspondence assembly + sigmoid + Kabsch for RegTR.forward (models/regtr.py:185-203):
 * kp (n,3) coarse key points, packed src clouds first then tgt clouds; corr (L,n,3) predicted
 * correspondences; logit (L,n) overlap logits; offs (2B+1) i32 cloud offsets.
 * pose (L,B,3,4): for pair b, a=[src_kp ; tgt_corr], b=[src_corr ; tgt_kp],
 * w=[sigmoid(src_logit) ; sigmoid(tgt_logit)]. */
int regtr_pose_from_corr(const float* kp, const float* corr, const float* logit,
                         const int32_t* offs, int n, int B, int L, float* pose, void* stream);

/* ---- status word helpers (device uint32) ------------------------------------------ */
int regtr_status_clear(uint32_t* status, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* REGTR_B200_H_ */
