/*
 * healswin_b200 -- C ABI of the B200-native HEAL-SWIN hot path.
 *
 * The reference (JanEGerken/HEAL-SWIN) has no FFI: its boundary for this path is
 * the Python module surface heal_swin.models_torch.{hp_windowing,hp_shifting,
 * swin_hp_transformer}.  This header is what a binding for that
 * boundary binds to; every entry point cites the reference code it replaces.
 * heal_swin_b200/_lib.py is the ctypes binding; INTEGRATION.md shows the
 * reference-side stub.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; the message is
 *     available from hs_last_error() (thread-local, never NULL).  Nothing calls
 *     exit()/abort().
 *   - "host" pointers are plain CPU memory, "dev" pointers are CUDA device memory of
 *     the current device.  All device entry points take the CUDA stream explicitly
 *     (cudaStream_t passed as void*), keep no mutable global state (only immutable
 *     caches: the driver entry point for TMA descriptors, the SM count) and are
 *     re-entrant (forward runs on the main thread, backward on the autograd thread).
 *   - tensors are dense row-major fp32 unless said otherwise.
 *   - dropout `seed` arguments: a plain 63-bit value, or -- bit 63 set -- an INDIRECT seed: bits 0-47 = address of a
 *     uint64 counter in device memory, bits 48-62 = a call id; the kernels mix the counter's current value with the id.
 *     A captured CUDA graph that increments the counter once per replay thereby draws fresh masks on every replay while
 *     the forward and the backward of one replay agree (heal_swin_b200/graph.py).
 */
#ifndef HEALSWIN_B200_H
#define HEALSWIN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HS_OK 0
#define HS_ERR_ARG 1      /* bad argument: the Python shim raises AssertionError  */
#define HS_ERR_CUDA 2     /* CUDA runtime / launch failure: RuntimeError          */
#define HS_ERR_UNSUPPORTED 3 /* shape outside what the kernels implement: RuntimeError */

/* shift strategies, swin_hp_transformer.py:277-304 */
#define HS_SHIFT_NONE 0       /* hp_shifting.NoShift           hp_shifting.py:31-39   */
#define HS_SHIFT_NEST_ROLL 1  /* hp_shifting.NestRollShift     hp_shifting.py:42-73   */
#define HS_SHIFT_NEST_GRID 2  /* hp_shifting.NestGridShift     hp_shifting.py:76-306  */
#define HS_SHIFT_RING 3       /* hp_shifting.RingShift         hp_shifting.py:309-404 */

/* operand precision of hs_gemm3 */
#define HS_GEMM_BF16X3 0
#define HS_GEMM_TF32 1
#define HS_GEMM_BF16 2

/* flags of the attention kernels */
#define HS_ATTN_COS 1u /* cosine attention, swin_hp_transformer.py:142-147 */
#define HS_ATTN_NO_TC 2u /* force the exact-fp32 CUDA-core kernels (cross-check of the tcgen05 TF32 path) */
#define HS_ATTN_NO_TRUNC_COMP 4u /* tcgen05 path: do not compensate the mean TF32 operand-truncation shrink (diagnostics) */
#define HS_MLP_GRAD16 256u /* hs_mlp_dgrad_gelu: z_dev holds g' = GELU'(z + b1) * dropmask as FP16 (hs_gemm3 mode 6), not z */

const char* hs_last_error(void);
int hs_version(void);
/* 0 when the current CUDA device is sm_100 (B200); fills name (may be NULL) */
int hs_device_info(int* sm_major, int* sm_minor, int* num_sms, char* name, int name_len);

/* ------------------------------------------------------------------ index layer (host) */

/* hp_windowing.get_nest_win_idcs, hp_windowing.py:43-62.  out: S*S int64, S = floor(sqrt(ws)) */
int hs_nest_win_idcs(int window_size, int64_t* out_host);
/* WindowAttention.relative_position_index buffer, swin_hp_transformer.py:98-114. out: ws*ws int64 */
int hs_rel_pos_index(int window_size, int64_t* out_host);
/* healpy.pixelfunc.nest2ring / ring2nest as called at hp_shifting.py:329,333 */
int hs_nest2ring(int64_t nside, const int64_t* in_host, int64_t* out_host, int64_t n);
int hs_ring2nest(int64_t nside, const int64_t* in_host, int64_t* out_host, int64_t n);
/*
 * Shifter tables for one block, swin_hp_transformer.py:271-308 + hp_shifting.py.
 *   shift_idcs[p]  : shifted[p] = x[shift_idcs[p]]        (NestGridShift/RingShift.shift_idcs; roll closed form)
 *   back_idcs[p]   : x'[p] = shifted'[back_idcs[p]]       (back_shift_idcs = argsort(shift_idcs))
 *   groups[p]      : mask group id of shifted pixel p     (get_mask(get_attn_mask=False))
 * N = base_pix * nside^2 entries each; any output may be NULL.  nside is the nside of the
 * *current* resolution (sqrt(input_resolution / base_pix), swin_hp_transformer.py:272).
 */
int hs_shift_tables(int strategy, int64_t nside, int base_pix, int window_size, int shift_size,
                    int64_t* shift_idcs_host, int64_t* back_idcs_host, int8_t* groups_host);
/* hp_shifting.get_attn_mask_from_mask, hp_shifting.py:10-28: (N,) ids -> (N/ws, ws, ws) in {0,-100} */
int hs_attn_mask_from_groups(const int8_t* groups_host, int64_t N, int window_size, float* mask_host);

/* ------------------------------------------------------------------ device kernels */

/*
 * Row gather out[b][p][:] = x[b][idx[p]][:], x/out (B, N, C) fp32, idx (N) int32: the standalone form
 * of shifter.shift (idx = shift_idcs) / shifter.shift_back (idx = back_shift_idcs),
 * hp_shifting.py:69-73, 302-306, 400-404.  Bit-exact (pure data movement).
 */
int hs_gather_rows(const float* x_dev, const int32_t* idx_dev, float* out_dev, int B, int64_t N, int C,
                   void* stream);

/*
 * Relative-position bias, swin_hp_transformer.py:152-159:
 *   bias[h][i][j] = table[index[i][j]][h]           table: (T, H) fp32, index: (ws*ws) int32, bias: (H, ws, ws)
 * and its adjoint dtable[t][h] += sum_{index[i][j]==t} dbias[h][i][j]  (dtable must be zero-filled or hold
 * the running gradient).
 */
int hs_rel_bias_expand(const float* table_dev, const int32_t* index_dev, float* bias_dev,
                       int T, int H, int ws, void* stream);
int hs_rel_bias_reduce(const float* dbias_dev, const int32_t* index_dev, float* dtable_dev,
                       int T, int H, int ws, void* stream);

/*
 * Windowed (shifted) multi-head self-attention core with the HEALPix shift, window
 * partition and window reverse folded into its loads and stores.  Replaces, in one kernel,
 * shifter.shift -> window_partition -> WindowAttention.forward[:142-171] -> window_reverse ->
 * shifter.shift_back (swin_hp_transformer.py:319-330, 136-171; hp_windowing.py:6-40;
 * hp_shifting.py:69-73, 302-306, 400-404).
 *
 *   qkv   (B, N, 3*C)   output of the qkv Linear on the UNSHIFTED token order; row layout [3][H][D]
 *   src   (N) int32     window w, slot j holds token src[w*ws + j] (shift_idcs); NULL = identity.
 *                       The output row of that slot is written back to the same token, which is
 *                       exactly shift_back o window_reverse.
 *   groups (N) uint8    mask group id per *shifted* slot, NULL = no mask.  logits get -100 where ids differ
 *   mask  (nW, ws, ws)  optional dense additive mask (WindowAttention.forward(x, mask) API), may be NULL
 *   bias  (H, ws, ws)   expanded relative position bias or NULL
 *   logit_scale (H)     raw parameter (cos attention: logits *= exp(min(ls, log 100))); else NULL
 *   scale               q scaling for the non-cos path (head_dim^-0.5 or qk_scale)
 *   out   (B, N, C)     attention output in the UNSHIFTED token order, row layout [H][D]
 *   lse   (P, H, B*N)   optional (may be NULL): forward statistics saved for the backward (which then needs no
 *                       statistics pass): plane 0 = per (head, token) log2-domain log-sum-exp of the logits row; with
 *                       HS_ATTN_COS P = 3 and planes 1, 2 = 1 / max(|q|, eps), 1 / max(|k|, eps); otherwise P = 1
 *   attn_drop, seed     dropout of the attention probabilities (nn.Dropout(attn_drop) at :167-169, training mode): entry
 *                       (i, j) of a (window, head) is zeroed with probability attn_drop and the rest scaled by
 *                       1 / (1 - attn_drop); the mask is a pure function of (seed, window, head, i, j), so the backward
 *                       must be given the same seed.  attn_drop = 0 disables it.
 */
int hs_window_attn_fwd(const float* qkv_dev, const int32_t* src_dev, const uint8_t* groups_dev,
                       const float* mask_dev, const float* bias_dev, const float* logit_scale_dev,
                       float scale, float attn_drop, uint64_t seed, float* out_dev, float* lse_dev, int B, int64_t N,
                       int C, int H, int ws, uint32_t flags, void* stream);
/*
 * Adjoint of hs_window_attn_fwd.  dqkv (B, N, 3C) is fully overwritten.  dbias (H, ws, ws) and
 * dlogit_scale (H) are accumulated into (+=), either may be NULL.  out / lse are the forward's outputs (the
 * tensor-core path uses them: rowsum(dO o O) and exp2(logit - lse) replace the statistics pass); with either NULL
 * the exact-fp32 kernels, which recompute everything from qkv, are used.
 */
int hs_window_attn_bwd(const float* qkv_dev, const float* out_dev, const float* lse_dev, const float* dout_dev,
                       const int32_t* src_dev,
                       const uint8_t* groups_dev, const float* mask_dev, const float* bias_dev,
                       const float* logit_scale_dev, float scale, float attn_drop, uint64_t seed, float* dqkv_dev,
                       float* dbias_dev, float* dlogit_scale_dev, int B, int64_t N, int C, int H, int ws,
                       uint32_t flags, void* stream);

/*
 * Row LayerNorm over the last dimension, optionally fused with the bias of the Linear that produced its input, the
 * dropout that follows that Linear, the per-sample stochastic-depth scale and the residual add of the v2 norm placement:
 *     y = residual + row_scale[row / rows_per_scale] * (LayerNorm(dropout(x + pre_bias)) * gamma + beta)
 * (pre_bias, residual, row_scale may be NULL; in_drop = 0 disables the dropout).
 * Replaces nn.LayerNorm at swin_hp_transformer.py:316, 333-338 (norm1 / norm2; "shortcut + drop_path(norm(proj_drop(
 * proj(.))))" and "x + drop_path(norm(mlp(.)))" are one pass, the proj / fc2 GEMMs run without bias), :392
 * (PatchMerging.norm, 4C), :428 (PatchExpand.norm, C/2), :450 (FinalPatchExpand_X4.norm), :781, :945.
 * x, residual, y: (rows, C) fp32; pre_bias, gamma, beta: (C); row_scale: (rows / rows_per_scale); mean, rstd: (rows) fp32
 * saved for the backward (both NULL = do not save).  Statistics as torch: biased variance, rstd = 1/sqrt(var + eps).
 * The dropout mask of element (row, col) is a pure function of (seed, row, col): pass the same seed to the backward.
 */
int hs_layernorm_fwd(const float* x_dev, const float* pre_bias_dev, const float* residual_dev, const float* gamma_dev,
                     const float* beta_dev, const float* row_scale_dev, int64_t rows_per_scale, float in_drop,
                     uint64_t seed, float* y_dev, float* mean_dev, float* rstd_dev, int64_t rows, int C, float eps,
                     void* stream);
/*
 * Adjoint: dx (rows, C) is overwritten; dgamma / dbeta / dpre_bias (C) are accumulated into (+=), each may be NULL
 * (dpre_bias = column sums of dx = the bias gradient of the producing Linear).  The gradient of the fused residual
 * input is dy itself.
 */
int hs_layernorm_bwd(const float* dy_dev, const float* x_dev, const float* pre_bias_dev, const float* mean_dev,
                     const float* rstd_dev, const float* gamma_dev, const float* row_scale_dev, int64_t rows_per_scale,
                     float in_drop, uint64_t seed, float* dx_dev, float* dgamma_dev, float* dbeta_dev,
                     float* dpre_bias_dev, int64_t rows, int C, void* stream);

/*
 * h = dropout(GELU(z + bias)) with the exact erf GELU of nn.GELU (swin_hp_transformer.py:21-44, Mlp: fc1 -> act -> drop),
 * z: (rows, C) the bias-free fc1 GEMM output, bias: (C) or NULL, drop = 0 disables the dropout (mask: a pure function of
 * (seed, row, col)); and its adjoint dz = dh * mask * GELU'(z + bias), dbias (C) += column sums of dz (may be NULL).
 * Covered: C % 4 == 0 and C <= 4096 (hs_bias_gelu_supported); other widths return HS_ERR_UNSUPPORTED.
 */
int hs_bias_gelu_supported(int64_t rows, int C); /* 1 when C % 4 == 0 and C <= 4096; else the caller uses torch ops */
int hs_bias_gelu_fwd(const float* z_dev, const float* bias_dev, float drop, uint64_t seed, float* h_dev, int64_t rows,
                     int C, void* stream);
int hs_bias_gelu_bwd(const float* dh_dev, const float* z_dev, const float* bias_dev, float drop, uint64_t seed,
                     float* dz_dev, float* dbias_dev, int64_t rows, int C, void* stream);

/*
 * Weight gradient of nn.Linear (autograd of F.linear at swin_hp_transformer.py:131, 172, 21-44, 394, 421, 444, 774):
 *     dw[n][k] += sum_t dy[t][n] * x[t][k]         dy: (T, N), x: (T, K), dw: (N, K), all fp32 row-major
 * and optionally the bias gradient in the same pass:   dbias[n] += sum_t dy[t][n]   (dbias may be NULL).
 * TF32 tensor-core kernel with the token range split over the SMs; dw / dbias are ACCUMULATED into (zero them for a
 * plain gradient).  hs_linear_wgrad_supported returns 0 when the shape is not covered (it then stays with the library
 * GEMM; covered: T >= 4096, N and K multiples of 4, max(N, K) >= 32, and min(N, K) <= 224 (any multiple of 4: a ragged
 * last 32-feature slab is zero-filled -- the patch embedding's 12 inputs), or a multiple of 32 and <= 512, or a multiple
 * of 64 and <= 1024 -- two launches over the column halves of the smaller operand), 1 when dw is covered, 2 when dbias
 * can be fused too (N >= K, K <= 224).  flags: HS_ATTN_NO_TRUNC_COMP only.
 */
int hs_linear_wgrad_supported(int64_t T, int N, int K);
int hs_linear_wgrad(const float* dy_dev, const float* x_dev, float* dw_dev, float* dbias_dev, int64_t T, int N, int K,
                    uint32_t flags, void* stream);

/*
 * Input gradient of the MLP's second half, fused (autograd of fc2(drop(GELU(fc1(x) + b1))) w.r.t. the fc1 output z,
 * swin_hp_transformer.py:21-44):
 *     dz[t][j] = (sum_c dy[t][c] * w2[c][j]) * GELU'(z[t][j] + b1[j]) * mask(seed, t, j)
 * dy: (T, C) gradient of the fc2 output, w2: (C, J) = fc2.weight, z: (T, J) fc1 output without bias, b1: (J) or NULL,
 * dz: (T, J); all fp32 row-major.  The (T, J) hidden gradient dy @ w2 lives only in tensor memory: this replaces the
 * fc2 input-gradient GEMM followed by hs_bias_gelu_bwd.  drop / seed: the mask hs_bias_gelu_fwd used.  TF32 tensor
 * cores.  hs_mlp_dgrad_gelu_supported returns 1 for covered shapes (C a multiple of 32 and <= 192, J a multiple of 128,
 * T >= 1024).  flags: HS_ATTN_NO_TRUNC_COMP; HS_MLP_GRAD16: z_dev is not z but the FP16 tensor g' = GELU'(z + b1) *
 * dropmask that hs_gemm3 mode 6 wrote in the forward -- then dz = (dy @ W2) * g' (b1, drop, seed are ignored), read straight
 * from global memory: a third fewer bytes and no GELU arithmetic in the backward.
 */
int hs_mlp_dgrad_gelu_supported(int64_t T, int C, int J);
int hs_mlp_dgrad_gelu(const float* dy_dev, const float* w2_dev, const float* z_dev, const float* b1_dev, float drop,
                      uint64_t seed, float* dz_dev, int64_t T, int C, int J, uint32_t flags, void* stream);

/*
 * Dense linear layers on the tensor cores at fp32-class accuracy (csrc/hs_gemm3_tc.cu): every fp32 operand is split
 * into two bf16 terms and the product accumulated in fp32 from three tcgen05 kind::f16 MMAs (hi*hi + lo*hi + hi*lo),
 * ~2^-16 relative per product -- this is what replaces the library (cuBLASLt TF32) GEMMs behind F.linear at
 * swin_hp_transformer.py:131-135 (qkv), :172 (proj), :21-44 (Mlp.fc1 / fc2), :394 (PatchMerging.reduction), :421
 * (PatchExpand.expand), :444 (FinalPatchExpand_X4.expand), :774 (concat_back_dim) and their input gradients.
 *
 * hs_weight_split prepares the weight operand once per optimizer step:
 *     out[r][c / 32][0:32] = bf16_hi(m(r, c)),  out[r][c / 32][32:64] = bf16_lo(m(r, c))      out: (rows, 2 * cols) bf16
 * with m(r, c) = w[r * ld + c] (transposed = 0: the forward operand of a (rows = N, cols = K) weight) or
 * m(r, c) = w[c * ld + r] (transposed = 1: the input-gradient operand, rows = K, cols = N, ld = K); the chunk count is
 * ceil(cols / 32) and the tail of a ragged last chunk is zero, i.e. out is (rows, 2 * ceil32(cols)).
 *
 * hs_gemm3:   acc[t][n] = sum_k a[t][k] * m(n, k)            a: (T, K) fp32, wsplit: (N, 2 ceil32(K)) bf16
 *   mode 0 (plain)      d = acc + bias
 *   mode 1 (add)        d = acc + bias + aux                  aux: (T, N) fp32, e.g. the residual-shortcut gradient
 *   mode 2 (gelu)       d = acc,  d2 = dropout(GELU(acc + bias))      Mlp fc1 + act + drop (:39-41); exact erf GELU
 *   mode 3 (gelu grad)  d = acc * GELU'(aux + bias) * dropmask        aux = the bias-free fc1 output z
 *   mode 6 (gelu, compact)  d2 = dropout(GELU(acc + bias)) as mode 2, but d receives g' = GELU'(acc + bias) * dropmask as
 *                       FP16 (d_dev is a (T, N) tensor of 2-byte elements): all the backward needs from the activation, at
 *                       a quarter of the bytes of z and h together; g' lies in [-0.13, 1.13] / (1 - drop); N % 32 == 0
 *   mode 7 (gelu grad, compact)  d = acc * g'                         aux_dev = that FP16 tensor (no bias, drop, seed)
 * (modes 4 and 5 are the LayerNorm forms below, with entry points of their own)
 * precision: HS_GEMM_BF16X3 (default: three bf16 MMAs per product, fp32-class) | HS_GEMM_TF32 (one TF32 MMA, A read as fp32
 * without conversion; wsplit must be hs_weight_split format 1 = fp32 rounded to the nearest TF32, same bytes per row: used
 * for the input gradients of the tensor-bound stages, where 5e-3 suffices) | HS_GEMM_BF16 (one bf16 MMA on the hi terms:
 * "bf16 operands, fp32 accumulate", BASELINE configs[3]; wsplit format 0).
 * colsum: (K) or NULL: colsum[k] += sum_t a[t][k] in the same pass -- when a is the output gradient of a linear (input-
 * gradient form) this is that linear's bias gradient, so no separate reduction over the activation is needed.
 * bias: (N) or NULL.  drop / seed: the element dropout of modes 2 / 3 (mask = pure function of (seed, row, col), as
 * hs_bias_gelu_fwd).  All of d, d2, aux are (T, N) fp32 row-major.  hs_gemm3_supported: N, K multiples of 4 (TMA row pitch);
 * ragged 32-wide chunks are zero-filled on load and clipped on store by the TMA unit.
 */
int hs_weight_split(const float* w_dev, int rows, int cols, int ld, int transposed, int format, uint16_t* out_dev,
                    void* stream);
/*
 * hs_weight_split for n matrices in one launch (the weight operands of a whole model, refreshed once per optimizer step).
 * descs_dev: n device-resident descriptors of 48 bytes each,
 *     { const float* w; uint16_t* out; int32 rows, cols, ld, transposed, format, tiles_x, tile0, pad; }
 * with tiles_x = ceil(cols / 32), tile0 = running sum of tiles_x * ceil(rows / 32) over the preceding descriptors
 * (ascending), total_tiles = that sum over all n.
 */
int hs_weight_split_batch(const void* descs_dev, int n, int total_tiles, void* stream);
int hs_gemm3_supported(int64_t T, int N, int K);
int hs_gemm3(const float* a_dev, const uint16_t* wsplit_dev, const float* bias_dev, const float* aux_dev, float* d_dev,
             float* d2_dev, float* colsum_dev, int64_t T, int N, int K, int mode, int precision, float drop, uint64_t seed,
             void* stream);

/*
 * Linear followed by LayerNorm over groups of G consecutive output columns, in the GEMM's epilogue (no pass of its own
 * over the (T, N) tensor):
 *     pre[t][n] = acc[t][n] + bias[n]
 *     y[t][n]   = (pre[t][n] - mean[t][n / G]) * rstd[t][n / G] * gamma[n % G] + beta[n % G]  (+ aux[t][n])
 * with mean / rstd the statistics of pre[t][g G : (g + 1) G] (biased variance, rstd = 1 / sqrt(var + eps)).
 *   G = N:      the `shortcut + norm(branch)` tail of a v2-placement block (swin_hp_transformer.py:333-338): the proj /
 *               Mlp.fc2 linear, its bias, norm1 / norm2 and the residual add (aux = shortcut) in one launch;
 *   G = N / 4:  PatchExpand (:420-430: Linear(C -> 2C) -> view (B, 4N, C/2) -> LayerNorm(C/2)); the (T, N) output IS the
 *               (4T, C/2) tensor of the view, and mean / rstd are laid out as its rows (index t * (N / G) + group).
 * pre_dev (the pre-norm tensor, what hs_layernorm_bwd needs), mean_dev / rstd_dev, bias_dev and aux_dev may be NULL.
 * hs_gemm3_ln_supported: hs_gemm3_supported and G a multiple of 32, G <= 192, N a multiple of G.
 * precision: HS_GEMM_BF16X3 or HS_GEMM_BF16.
 */
int hs_gemm3_ln_supported(int64_t T, int N, int K, int G);
int hs_gemm3_ln(const float* a_dev, const uint16_t* wsplit_dev, const float* bias_dev, const float* gamma_dev,
                const float* beta_dev, const float* aux_dev, float* pre_dev, float* y_dev, float* mean_dev, float* rstd_dev,
                int64_t T, int N, int K, int G, float eps, int precision, void* stream);

/*
 * LayerNorm over the whole input row followed by a bias-free Linear, without materialising the normalised tensor:
 * PatchMerging (swin_hp_transformer.py:378-395: view (B, N/4, 4C) -> LayerNorm(4C) -> Linear(4C -> 2C); in NESTED order the
 * four siblings of a token are consecutive, so the gather IS the view).  With W' = W diag(gamma):
 *     d[t][n] = rstd[t] * (sum_k a[t][k] W'[n][k] - mean[t] * wsum[n]) + b0[n],   wsum = W' 1,  b0 = W beta
 * wsplit: hs_weight_split of W' (format 0).  The GEMM runs on the raw rows; its operand converters take the row statistics
 * (exact fp32, pivoted sums) and the epilogue applies them.  mean_dev / rstd_dev (T; what hs_layernorm_bwd needs) and
 * b0_dev may be NULL.  hs_gemm3_lnin_supported: hs_gemm3_supported and K a multiple of 32, K >= 160.
 */
int hs_gemm3_lnin_supported(int64_t T, int N, int K);
int hs_gemm3_lnin(const float* a_dev, const uint16_t* wsplit_dev, const float* wsum_dev, const float* b0_dev, float* d_dev,
                  float* mean_dev, float* rstd_dev, int64_t T, int N, int K, float eps, int precision, void* stream);

/*
 * Decoder tail, fused: logits = Conv1d_1x1(LayerNorm(x)) (FinalPatchExpand_X4.norm + SwinHPTransformerSys.output,
 * swin_hp_transformer.py:450, 781-786, 945) in one pass, and its backward in one pass.
 *   x: (rows, C) fp32 with rows = B * rows_per_sample; gamma, beta: (C); w: (K, C) = output.weight[:, :, 0];
 *   head_bias: (K) or NULL; logits / dlogits: (B, K, rows_per_sample) -- the layout of the reference's network output;
 *   mean, rstd: (rows) LayerNorm statistics, written by the forward and read by the backward.
 * Backward writes dx (rows, C) and ACCUMULATES  s_acc[k][c] += sum_rows dlogits[row][k] * xhat[row][c]  (K, C) and
 * g_acc[k] += sum_rows dlogits[row][k]  (K) (zero them first), from which
 *   d(w) = gamma * s_acc + beta * g_acc,  d(gamma) = sum_k w * s_acc,  d(beta) = sum_k w * g_acc,  d(head_bias) = g_acc.
 * hs_ln_head_supported: 1 for C in {32, 64, 96, 128} and 1 <= K <= 16.
 */
int hs_ln_head_supported(int64_t rows, int C, int K);
int hs_ln_head_fwd(const float* x_dev, const float* gamma_dev, const float* beta_dev, const float* w_dev,
                   const float* head_bias_dev, float* logits_dev, float* mean_dev, float* rstd_dev, int64_t rows,
                   int64_t rows_per_sample, int C, int K, float eps, void* stream);
int hs_ln_head_bwd(const float* dlogits_dev, const float* x_dev, const float* mean_dev, const float* rstd_dev,
                   const float* gamma_dev, const float* w_dev, float* dx_dev, float* s_acc_dev, float* g_acc_dev,
                   int64_t rows, int64_t rows_per_sample, int C, int K, void* stream);

/*
 * Cross-entropy over the network output, forward and gradient in one pass (nn.CrossEntropyLoss,
 * heal_swin/models_lightning/segmentation/model_lightning_swin_hp.py:45, 109):
 *   logits, dlogits: (B, K, P) fp32 (class planes of P pixels, the layout of the network output); target: (B, P) uint8 or
 *   int64 (target_bytes = 1 | 8); pixels whose target equals ignore_index (or lies outside [0, K)) are not counted.
 *   dlogits = softmax(logits) - onehot(target) (0 for uncounted pixels) -- UNSCALED; acc[0] += sum of -log softmax[target],
 *   acc[1] += number of counted pixels (zero acc first).  mean loss = acc[0] / acc[1]; d(mean loss) = dlogits / acc[1].
 * hs_cross_entropy_supported: 1 <= K <= 32.
 */
int hs_cross_entropy_supported(int K);
int hs_cross_entropy(const float* logits_dev, const void* target_dev, int target_bytes, float* dlogits_dev, float* acc_dev,
                     int B, int K, int64_t P, int64_t ignore_index, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HEALSWIN_B200_H */
