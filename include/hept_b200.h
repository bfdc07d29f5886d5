/* hept_b200 — C ABI of the sm_100a library behind the HEPT attention drop-in.
 *
 * The reference (Graph-COM/HEPT) is pure PyTorch: it has no FFI of its own.  The boundary it
 * offers is the Python module `HEPTAttention.forward(query, key, value, **kwargs)`
 * (example/hept.py:43-81, src/models/attention/hept.py:71-117).  These entry points are what a
 * binding for that path calls; each one names the reference lines it replaces.  The host-side
 * mirror of the module (same ctor / forward / state_dict) is hept_b200/attention.py, which binds
 * this header through ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless named `*_host`; tensors are dense row-major and 16-byte aligned (rows are
 *     read and written as 16-byte vectors; out_linear and the Attn front check it and return HEPT_EINVAL);
 *   - the library never allocates, frees or synchronises: the caller (torch) owns every buffer,
 *     including `workspace`, and passes the CUDA stream to enqueue on (`stream` is a cudaStream_t
 *     passed as void*);
 *   - every function returns 0 on success or a negative HEPT_E* code; hept_last_error() returns
 *     a thread-local message for the last failure;
 *   - symbols: N hits (multiple of B), H heads, D dims/head, C coords_dim, E = D + C,
 *     T n_hashes, B block_size, R = C - 1, K num_w_per_dist.
 *
 * Supported compile-time shapes: D = 24 with C in {6 (tracking), 4 (pileup)} and B in {64, 100, 128}, and (D, C, B) =
 * (8, 6, 10) for tests.  Anything else returns HEPT_EUNSUPPORTED (no slow fallback, no CPU path); T <= 4, H <= 32.
 * Blocks of 128 hits do not fit the TMEM layout of the tcgen05 tiles and run on the fp32 CUDA-core tiles.
 * Devices: kernels run on the CURRENT CUDA device; the caller makes the device that owns the pointers current
 * (hept_b200/ops.py does so around every call).  Per-device state (shared-memory opt-ins, SM counts) is kept per device.
 */
#ifndef HEPT_B200_H
#define HEPT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HEPT_OK 0
#define HEPT_EINVAL (-1)       /* bad argument (null pointer, N % B != 0, ...) */
#define HEPT_EUNSUPPORTED (-2) /* (D, C, B) combination not compiled in */
#define HEPT_ECUDA (-3)        /* a CUDA launch failed; see hept_last_error() */
#define HEPT_EWORKSPACE (-4)   /* workspace too small */

/* Problem shape shared by all calls. raw_size: rows >= raw_size are the zero/+inf padding of the
 * src/ flavour (src/models/attention/hept.py:89-96); pass raw_size == N for the example/ flavour. */
typedef struct hept_shape {
  int32_t N, H, D, C, T, B;
  int32_t raw_size;
} hept_shape;

int hept_abi_version(void);
const char* hept_last_error(void);
/* 1 if kernels for (D, C, B) are compiled in. */
int hept_shape_supported(int32_t D, int32_t C, int32_t B);

/* ---- a3  prep_qk coordinate scale (example/hept.py:21-25) -------------------------------------
 * scale[h,c] = sqrt(2 * qw[h, max(c-1,0)]),  qw[h,r] = sum_k exp(min(sum_d w[h*D+d, r*K+k], 50)).
 * w_rpe_weight (H*D, R*K) -> scale (H, C=R+1). */
int hept_coord_scale_fwd(const float* w_rpe_weight, int32_t H, int32_t D, int32_t R, int32_t K,
                         float* scale, void* stream);
/* gradient of the above: dscale (H, C) -> dw (H*D, R*K) (autograd of example/hept.py:22-25). */
int hept_coord_scale_bwd(const float* w_rpe_weight, const float* scale, const float* dscale, int32_t H,
                         int32_t D, int32_t R, int32_t K, float* dw, void* stream);

/* ---- a4+a5  E2LSH.forward x2 + lsh_mapping (example/hept_utils.py:45-47, 64-71) ---------------
 * q,k (N, H*D); coords (N, C); scale (H, C); alpha (H, E, T).
 * proj (2, T, H, N): [0] = queries, [1] = keys.  span (T, H) = max - min over both.
 * workspace: hept_hash_workspace_bytes(s) = one (min, max) pair per (CTA, table, head), at most HEPT_HASH_MAX_CTAS CTAs
 * (the persistent grid keeps extrema in registers; no atomics, so the span is deterministic). */
#define HEPT_HASH_MAX_CTAS 1024
size_t hept_hash_workspace_bytes(const hept_shape* s);
int hept_hash_project(const hept_shape* s, const float* q, const float* k, const float* coords,
                      const float* scale, const float* alpha, float* proj, float* span,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ---- a3, coordinate half of prep_qk (example/hept.py:25): hat_coords (N, H, 8) = scale[h,c] * coords[n,c], zero
 * padded (rows >= raw_size zero).  Shared by every table and by the query / key side of the tile kernels. */
int hept_hat_coords(const hept_shape* s, const float* coords, const float* scale, float* hat_coords, void* stream);

/* ---- a6  AND-construction, example/ flavour (example/hept.py:63-65) ---------------------------
 * keys[s,t,h,n] = proj[s,t,h,n] + float(combined_shifts[t,h,n]) * span[t,h]   (convert, mul, add). */
int hept_keys_from_packed_shifts(const hept_shape* s, const float* proj, const float* span,
                                 const int64_t* combined_shifts, float* keys, void* stream);
/* the same with the codes as int32 (hept_prepare_batched emits both; the values are identical, 96 bytes per hit fewer) */
int hept_keys_from_packed_shifts32(const hept_shape* s, const float* proj, const float* span,
                                   const int32_t* combined_shifts32, float* keys, void* stream);
/* ---- a6' AND-construction, src/ flavour (src/models/attention/hept.py:46-56, 93-101) ----------
 * region_eta/phi (T*H, N) float, regions_h (2, T*H); rows >= raw_size get +inf. */
int hept_keys_from_region_indices(const hept_shape* s, const float* proj, const float* span,
                                  const float* region_eta, const float* region_phi,
                                  const float* regions_h, float* keys, void* stream);

/* ---- a7  argsort x2 (example/hept.py:67-68), tie-break = stable ascending ---------------------
 * keys (num_segments, n) float32 -> positions (num_segments, n) int32, one stable LSD radix sort
 * per segment.  -0.0 sorts equal to +0.0; NaN sorts last. */
size_t hept_argsort_workspace_bytes(int32_t num_segments, int32_t n);
int hept_segmented_argsort(const float* keys, int32_t num_segments, int32_t n, int32_t* positions,
                           void* workspace, size_t workspace_bytes, void* stream);

/* ---- a8+a9+a10+a11  sort_to_buckets, qkv_res, unsort_from_buckets ------------------------------
 * (example/hept_utils.py:74-97, example/hept.py:7-18,70-78).  positions (2, T, H, N) int32.
 * hat_coords (N, H, 8) from hept_hat_coords (required by the tcgen05 engine, may be null for the fp32 engine).
 * Output `stage` (H, N, T, 32): per hit/head/table one 128-byte row holding the numerator
 * so[0..D) and, at [D], the normaliser denom = rowsum + 1e-20, already back in ORIGINAL hit order. */
int hept_block_attention_fwd(const hept_shape* s, const float* q, const float* k, const float* v,
                             const float* coords, const float* scale, const float* hat_coords,
                             const int32_t* positions, float* stage, void* stream);

/* ---- a12  OR-combine (example/hept.py:79) -----------------------------------------------------
 * out_pre (N, H*D) = sum_t numer / sum_t denom ; den_sum (N, H) kept for backward. */
int hept_or_combine(const hept_shape* s, const float* stage, float* out_pre, float* den_sum, void* stream);

/* ---- a12, second half: out_linear (example/hept.py:80; nn.Linear(H*D, D)) ------------------------
 * out (N, D) = out_pre (N, H*D) weight^T (D, H*D) + bias (D), no library GEMM.  Arithmetic: for (H, D) = (8, 24) the three
 * products of this section run as 3xTF32 tensor-core products (every operand split into tf32 hi + lo, hi*hi + hi*lo + lo*hi
 * accumulated in fp32: ~2^-22 relative per product, like the attention tiles); other shapes use fp32 FMAs. */
int hept_out_linear_fwd(const hept_shape* s, const float* out_pre, const float* weight, const float* bias,
                        float* out, void* stream);
/* its backward: d_out (N, D) -> d_out_pre (N, H*D) = d_out weight (skipped when d_out_pre is null),
 * d_weight (D, H*D) = d_out^T out_pre, d_bias (D) = column sums of d_out.  The parameter gradients are summed per
 * CTA over slabs of hits, then over at most HEPT_OUT_LINEAR_MAX_CTAS CTAs in a fixed order: deterministic. */
#define HEPT_OUT_LINEAR_MAX_CTAS 1024
size_t hept_out_linear_bwd_workspace_bytes(const hept_shape* s);
int hept_out_linear_bwd(const hept_shape* s, const float* d_out, const float* weight, const float* out_pre,
                        float* d_out_pre, float* d_weight, float* d_bias, void* workspace,
                        size_t workspace_bytes, void* stream);

/* ---- a18  backward of a8..a12 and of the coordinate scale's use in a3 --------------------------
 * d_out_pre (N, H*D) is the gradient arriving from out_linear.  Writes dq, dk, dv (N, H*D) and
 * dscale (H, C).  Deterministic (no floating-point atomics). */
size_t hept_attention_bwd_workspace_bytes(const hept_shape* s);
int hept_block_attention_bwd(const hept_shape* s, const float* q, const float* k, const float* v,
                             const float* coords, const float* scale, const int32_t* positions,
                             const float* out_pre, const float* den_sum, const float* d_out_pre,
                             float* dq, float* dk, float* dv, float* dscale, void* workspace,
                             size_t workspace_bytes, void* stream);

/* ---- whole forward of a3..a12 in one call (fewer host round trips) ----------------------------
 * Exactly one of combined_shifts / (region_eta, region_phi, regions_h) is non-null.
 * Outputs: scale (H,C), positions (2,T,H,N), out_pre (N,H*D), den_sum (N,H). */
size_t hept_attention_fwd_workspace_bytes(const hept_shape* s);
int hept_attention_fwd(const hept_shape* s, const float* q, const float* k, const float* v,
                       const float* coords, const float* w_rpe_weight, int32_t K, const float* alpha,
                       const int64_t* combined_shifts, const float* region_eta, const float* region_phi,
                       const float* regions_h, float* scale, int32_t* positions, float* out_pre,
                       float* den_sum, void* workspace, size_t workspace_bytes, void* stream);

/* ---- a13..a17  per-forward preparation of the AND-hash inputs -----------------------------------------------------
 * example/ flavour: prepare_input of example/transformer.py:35-63 (quantile_partition example/hept_utils.py:6-14, bit_shift
 * example/transformer.py:10-13, pad_and_unpad :16-32) for a batch of events, without the reference's per-event host loop.
 *   coords (n_raw, C) fp32 (columns 0 / 1 = eta / phi); batch (n_raw) int64, ascending; event_start, pad_start: DEVICE
 *   arrays of num_events + 1 int32 offsets of the events in raw / padded order (the caller knows the event sizes: they
 *   decide n_pad, the size of every output); max_event = the longest event; regions_h (2, TH) fp32 = the model's `regions`
 *   parameter rearranged "c a h -> a (c h)"; block_size B; code_bits = an upper bound on the number of bits of the
 *   (table 0, head 0) code, ceil(log2(num_events)) + 2 ceil(log2(max regions per axis + 2)) (0 = unknown, 32): the padding
 *   order is a radix sort of those integer codes, one pass per 8 bits.
 * Outputs, all in padded order: combined_shifts (TH, n_pad) int64 (+ the same values as int32 when combined_shifts32 is
 * non-null), take (n_pad) int64 = index of the raw point every padded row shows (the reference's pad_seq: x[take]),
 * is_real (n_pad) uint8 (the reference's unpad_seq), coords_pad (n_pad, C).
 * Sort tie-break: stable (the reference's argsort is not; only points with EQUAL eta, phi or code can differ). */
size_t hept_prepare_batched_workspace_bytes(int32_t n_raw, int32_t num_events, int32_t max_event);
int hept_prepare_batched(const float* coords, int32_t C, const int64_t* batch, const int32_t* event_start,
                         const int32_t* pad_start, int32_t num_events, int32_t n_raw, int32_t n_pad, int32_t max_event,
                         const float* regions_h, int32_t TH, int32_t block_size, int32_t code_bits, int64_t* combined_shifts,
                         int32_t* combined_shifts32, int64_t* take, uint8_t* is_real, float* coords_pad, void* workspace,
                         size_t workspace_bytes, void* stream);
/* src/ flavour: HEPT branch of prepare_input, src/models/baselines/transformer.py:43-57 (one event): coords padded with +inf
 * to n_pad rows for the ranking, region_eta / region_phi (TH, n_pad) fp32 = quantile regions of the eta / phi ranks, then
 * the padding rows of coords_pad (n_pad, C) set to zero. */
size_t hept_prepare_single_workspace_bytes(int32_t n_pad);
int hept_prepare_single(const float* coords, int32_t C, int32_t n_raw, int32_t n_pad, const float* regions_h, int32_t TH,
                        float* coords_pad, float* region_eta, float* region_phi, void* workspace, size_t workspace_bytes,
                        void* stream);

/* hept_attention_fwd for the example/ flavour with int32 codes */
int hept_attention_fwd_shifts32(const hept_shape* s, const float* q, const float* k, const float* v, const float* coords,
                                const float* w_rpe_weight, int32_t K, const float* alpha, const int32_t* combined_shifts32,
                                float* scale, int32_t* positions, float* out_pre, float* den_sum, void* workspace,
                                size_t workspace_bytes, void* stream);

/* ---- the caller's other LayerNorms: the Attn block's norm2 (example/transformer.py:163) and the model head's 256-wide norms
 * (torch_geometric MLP, example/transformer.py:84).  torch.nn.functional.layer_norm over the last dimension: biased variance,
 * eps inside the square root, y = (x - mean) rstd weight + bias; fp32.  4 <= D <= 256, D % 4 == 0.
 *   forward:  x (N, D) -> y (N, D), mean_rstd (N, 2) kept for the backward
 *   backward: dy (N, D) -> dx (N, D), d_weight, d_bias (D); deterministic (fixed-order reductions). */
int hept_layer_norm_supported(int32_t D);
int hept_layer_norm_fwd(const float* x, const float* weight, const float* bias, int32_t N, int32_t D, float eps, float* y,
                        float* mean_rstd, void* stream);
size_t hept_layer_norm_bwd_workspace_bytes(int32_t N, int32_t D);
int hept_layer_norm_bwd(const float* x, const float* mean_rstd, const float* weight, const float* dy, int32_t N, int32_t D,
                        float* dx, float* d_weight, float* d_bias, void* workspace, size_t workspace_bytes, void* stream);

/* ---- the front of the caller's Attn block ------------------------------------------------------
 * x_normed = norm1(x); q, k, v = w_q(x_normed), w_k(x_normed), w_v(x_normed)   (example/transformer.py:157-158,
 * src/models/baselines/transformer.py:209-212): LayerNorm over D (eps inside the square root) and three bias-free
 * Linear(D, H*D).  With it a hit crosses the boundary as its D-float activation row instead of three H*D-float rows.
 *   x (N, D); norm_weight, norm_bias (D); w_q, w_k, w_v (H*D, D) as nn.Linear stores them.
 *   forward outputs: wt (3, D, H*D) the transposed weights, x_normed (N, D) (both kept for the backward), q, k, v (N, H*D).
 *   backward: dq, dk, dv (N, H*D) -> dx (N, D) (the gradient through norm1 only: the block's residual path is the caller's),
 *   d_norm_weight, d_norm_bias (D), d_w_q, d_w_k, d_w_v (H*D, D); deterministic (fixed-order reductions).
 * The products are 3xTF32 tensor-core products with fp32 accumulation (see out_linear above); LayerNorm and its backward are fp32.
 * Compiled for (H, D) = (8, 24): hept_attn_qkv_supported says so; other shapes return HEPT_EUNSUPPORTED. */
int hept_attn_qkv_supported(int32_t H, int32_t D);
int hept_attn_qkv_fwd(const float* x, const float* norm_weight, const float* norm_bias, const float* w_q, const float* w_k,
                      const float* w_v, int32_t N, int32_t H, int32_t D, float eps, float* wt, float* x_normed, float* q,
                      float* k, float* v, void* stream);
size_t hept_attn_qkv_bwd_workspace_bytes(int32_t N, int32_t H, int32_t D);
int hept_attn_qkv_bwd(const float* x, const float* x_normed, const float* norm_weight, const float* wt, const float* dq,
                      const float* dk, const float* dv, int32_t N, int32_t H, int32_t D, float eps, float* dx,
                      float* d_norm_weight, float* d_norm_bias, float* d_w_q, float* d_w_k, float* d_w_v, void* workspace,
                      size_t workspace_bytes, void* stream);

/* ---- SURVEY.md 8(f)-4, first half: InfoNCE loss of the tracking task (src/utils/losses.py:8-74) ---------------------------
 * x (N, d <= 16) embeddings; point_pairs (2, P) int64; cluster_ids (N) int64; recons, pts (N) fp32; metric 0 = l2_rbf,
 * 1 = l2_inverse, 2 = cosine (losses.py:20-30); pt_thres 0.9 in the reference (losses.py:17).  loss: one device float.
 * `saved` (hept_infonce_saved_bytes) carries the forward's state to the backward (scores, pair flags, the CSR index of the
 * pairs by first point, per-point denominators, label sizes); `workspace` is scratch (hept_infonce_workspace_bytes).
 * Deterministic: every floating-point sum runs in a fixed order (no floating-point atomics).  grad_loss: one device float. */
size_t hept_infonce_saved_bytes(int32_t N, int64_t P);
size_t hept_infonce_workspace_bytes(int32_t N, int64_t P, int32_t backward);
int hept_infonce_fwd(const float* x, int32_t N, int32_t d, const int64_t* point_pairs, int64_t P, const int64_t* cluster_ids,
                     const float* recons, const float* pts, float pt_thres, int32_t metric, float tau, float* loss, void* saved,
                     size_t saved_bytes, void* workspace, size_t workspace_bytes, void* stream);
int hept_infonce_bwd(const float* x, int32_t N, int32_t d, const int64_t* point_pairs, int64_t P, int32_t metric, float tau,
                     const float* grad_loss, const void* saved, size_t saved_bytes, float* dx, void* workspace,
                     size_t workspace_bytes, void* stream);

/* ---- SURVEY.md 8(f)-4, second half: kNN metrics (acc_and_pr_at_k + calc_scores, src/utils/metrics.py:23-93) --------------
 * x (N, d <= 16) embeddings, cluster_ids (N) int64, queries (M) int64 = the points that pass the reference's mask, in order;
 * cosine != 0: distance 1 - cos instead of Euclidean; K neighbours after dropping the nearest (the query itself).
 * out (5 floats, device): mean accuracy, precision, recall over the queries whose cluster has more than one point, their
 * number, and the largest k = cluster size - 1 seen (the reference asserts it is <= K).  Deterministic. */
size_t hept_knn_metrics_workspace_bytes(int32_t N, int32_t M);
int hept_knn_metrics(const float* x, int32_t N, int32_t d, const int64_t* cluster_ids, const int64_t* queries, int32_t M,
                     int32_t cosine, int32_t K, float* out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- the gradient all-reduce of the data-parallel training step as ONE kernel over NVLink peer memory -----
 * bufs_dev: device array of `world` base pointers of the ranks' symmetric buffers, each laid out [hept_p2p_flag_bytes() of
 * flags, zero before the first call | n_floats of data]; n_floats % 4 == 0; seq: a call counter, the same on every rank,
 * increasing by 1 per call, never 0.  Every rank's data becomes scale * (sum over ranks, in rank order: the same bits on every
 * rank).  scratch: n_floats of local memory.  *err is set non-zero if a peer does not arrive within ~1 s (the kernel never
 * hangs).  The mapping of the peers' buffers is the host's business (hept_b200/sharding.py uses torch symmetric memory). */
size_t hept_p2p_flag_bytes(void);
int hept_p2p_allreduce(void* const* bufs_dev, int32_t rank, int32_t world, int64_t n_floats, uint32_t seq, float scale,
                       float* scratch, int32_t* err, void* stream);

/* kernel launches this library enqueued (any thread of the process) since the counter was last reset
 * (bench.py's gpu_launches); reset != 0 zeroes the counter after reading it. */
int hept_launch_count(int reset);

/* profiling aid for bench.py / ncu: choose which kernels hept_block_attention_bwd launches
 * (bit 0: dq tile kernel, bit 1: dk/dv tile kernel, bit 2: table reduction); default 7 = all. */
void hept_set_bwd_stage_mask(int mask);

/* tile engine for a8-a11: 0 = fp32 CUDA-core tiles (attn_fwd.cu), 1 = tcgen05 tensor-core tiles with
 * 3xTF32 operand splitting (attn_fwd_tc.cu).  Process-wide; both engines meet the same parity tolerance. */
void hept_set_engine(int engine);
int hept_get_engine(void);

/* a7 segmented argsort: 0 = cluster-resident sort (one thread-block cluster per segment, passes through distributed
 * shared memory) whenever a segment fits, global passes otherwise (default); 1 = global passes always.  Process-wide;
 * both produce the same positions. */
void hept_set_sort_variant(int variant);
int hept_get_sort_variant(void);

/* backward tile kernels: 1 = fp32 CUDA-core tiles, one lane per row (attn_bwd.cu), 3 = tcgen05 tiles (attn_bwd_tc.cu;
 * default).  The tcgen05 tiles either add the T tables' rows straight into dq / dk / dv in table order (no staging rows,
 * no summing kernel; launched cooperatively, because its CTAs wait for each other) or stage them per table: 3 picks the
 * direct form when a (head, table) group is at least two waves of tiles, 4 = direct whenever possible, 5 = staged always.
 * Same bits either way.  Process-wide test / profiling aid (atomic). */
void hept_set_bwd_variant(int variant);
int hept_get_bwd_variant(void);

#ifdef __cplusplus
}
#endif
#endif /* HEPT_B200_H */
