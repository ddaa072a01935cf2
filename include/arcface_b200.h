/*
 * arcface_b200.h -- C ABI of libarcface_b200.so: the B200 (sm_100a) implementation of the ArcFace
 * additive-angular-margin head + softmax cross-entropy (forward, argmax, backward).
 *
 * Reference path replaced (forrestsocool/MultimodalSimilar, read-only at /root/reference):
 *   arcface.py:45-63   ArcMarginProduct.forward        (normalise, cosine GEMM, margin, one-hot blend, scale)
 *   arcface.py:65-67   ArcMarginProduct.forward_test   (bare cosines)
 *   nlp_classifier_train.py:100,120-123  nn.CrossEntropyLoss()(preds, y), loss.backward(), torch.argmax
 * The reference has no FFI of its own (it is pure PyTorch); its "plugin API" is the nn.Module protocol of
 * arcface.ArcMarginProduct.  multimodalsimilar_b200/head.py mirrors that protocol and binds the entry
 * points below with ctypes; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch / CUDA types (a CUDA stream is passed as void*).
 *   - every device buffer is owned by the caller (the library allocates nothing persistent);
 *     pointers must be 16-byte aligned and rows contiguous unless a leading dimension is given.
 *   - all work is enqueued asynchronously on `stream`; no entry point synchronises the device,
 *     except the *_host convenience call which waits for its own results.
 *   - return value: ARCFACE_B200_OK (0) or a negative ARCFACE_B200_E_* code; the message is available
 *     from arcface_b200_last_error() (thread-local).  Nothing throws or aborts.
 *   - there is no CPU fallback: on a device that is not compute capability 10.x every compute entry
 *     point returns ARCFACE_B200_E_ARCH.
 *   - bf16 buffers are passed as uint16_t*.
 *   - shapes: B batch rows (1..ARCFACE_B200_MAX_BATCH), D embedding width (D % 8 == 0), C_local classes held by this rank
 *     (the whole C on one GPU), class_offset = first global class id of this rank's shard.
 */
#ifndef ARCFACE_B200_H_
#define ARCFACE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ARCFACE_B200_VERSION_MAJOR 0
#define ARCFACE_B200_VERSION_MINOR 1

#define ARCFACE_B200_OK 0
#define ARCFACE_B200_E_ARCH (-1)      /* device is not sm_100 */
#define ARCFACE_B200_E_SHAPE (-2)     /* unsupported B / D / C */
#define ARCFACE_B200_E_LAYOUT (-3)    /* misaligned pointer or bad leading dimension */
#define ARCFACE_B200_E_WORKSPACE (-4) /* workspace too small */
#define ARCFACE_B200_E_CUDA (-5)      /* a CUDA runtime / driver call failed */
#define ARCFACE_B200_E_ARG (-6)       /* null pointer or invalid scalar */

/* Precision modes of the cosine contraction (the `precision` keyword of the module).
 *   BF16   : bf16 operands, fp32 accumulation -- the throughput mode (logits within ~2e-2 * s / 30 of fp32).
 *   BF16X3 : every normalised value is a bf16 pair hi + lo and the contraction keeps hi.hi + hi.lo + lo.hi
 *            (three tensor-core products in one fp32 accumulator): cosines within 1e-5 of fp32, i.e. better than
 *            the "1e-4 under a TF32 mode" of the north star, on the same tcgen05 kind::f16 kernels. */
#define ARCFACE_B200_PREC_BF16 0
#define ARCFACE_B200_PREC_BF16X3 1

#define ARCFACE_B200_MAX_BATCH 2048

int32_t arcface_b200_version(int32_t* major, int32_t* minor);
const char* arcface_b200_last_error(void);
/* ARCFACE_B200_OK iff the current CUDA device can run the kernels (compute capability 10.x). */
int32_t arcface_b200_device_ok(void);

/* K1 -- fused row L2-normalise + bf16 cast.  Replaces F.normalize(x) / F.normalize(self.weight)
 * (arcface.py:47).  dst[r, :] = bf16(src[r, :] / max(||src[r, :]||, 1e-12)), inv_norm[r] = 1 / max(||.||, 1e-12).
 * dst_t (nullable) additionally receives the transpose, dst_t[d * ld_t + r], for the dW GEMM. */
int32_t arcface_b200_normalize_cast(const float* src, int64_t rows, int32_t D, uint16_t* dst, float* inv_norm,
                                    uint16_t* dst_t, int64_t ld_t, void* stream);

/* K1 of the ARCFACE_B200_PREC_BF16X3 mode: each normalised value v as the bf16 pair hi = bf16(v), lo = bf16(v - hi),
 * the row written three times along the contraction dimension -- dst3 [rows][3 D]: order 0 (embeddings) [hi|hi|lo],
 * order 1 (class weights) [hi|lo|hi] -- so that the bf16 GEMMs over 3 D columns (arcface_b200_forward_stats /
 * _logits / _cosine_topk with D := 3 D) accumulate hi.hi + hi.lo + lo.hi in fp32.  dst_t (nullable, [D][3 ld_t], caller
 * zeroes the padding): the transposed operand of the dW GEMM, [hi^T | lo^T | hi^T]. */
int32_t arcface_b200_normalize_cast3(const float* src, int64_t rows, int32_t D, int32_t order, uint16_t* dst3,
                                     float* inv_norm, uint16_t* dst_t, int64_t ld_t, void* stream);

/* Class sampling (PartialFC-style; SURVEY.md section 8f row N4 -- the reference always trains all classes, the rule
 * restated in oracle/arcface_numpy.py:partial_fc_sample is insightface's partial_fc_v2 `sample`).
 * normalize_cast_gather: K1 over the sampled rows, dst[r, :] = bf16 normalise of src[index[r], :] (index sorted or not,
 * values in [0, src_rows)); the forward / backward kernels then run on the `rows`-class sub-matrix.
 * scatter_rows: dst[index[r], :] = src[r, :] -- the sampled rows' gradient back into the full-size dW (the caller
 * zeroes the rest). */
int32_t arcface_b200_normalize_cast_gather(const float* src, int64_t src_rows, const int64_t* index, int64_t rows,
                                           int32_t D, uint16_t* dst, float* inv_norm, void* stream);
int32_t arcface_b200_scatter_rows(const float* src, const int64_t* index, int64_t rows, int32_t D, float* dst,
                                  int64_t dst_rows, void* stream);

/* dst[i] += src[i], i < n (fp32; n % 4 == 0, 16-byte aligned).  A global batch above 1024 rows runs the GEMM kernels once
 * per chunk of rows (the row statistics and the exchanges stay one pass); this sums the chunks' dW. */
int32_t arcface_b200_accumulate(float* dst, const float* src, int64_t n, void* stream);

/* Label column in fp32 + margin (arcface.py:49-55 restricted to the label column, the only place the
 * reference's one-hot blend at :58-60 uses phi).  For every row b whose label falls in this shard:
 *   t = <x_b, w_y> * inv_nx[b] * inv_nw[y];  sine = sqrt(max(0, 1 - t^2));  phi = t cos_m - sine sin_m
 *   u = easy ? (t > 0 ? phi : t) : (t - th > 0 ? phi : t - mm);   z_label = s * u
 *   dphi = d u / d t   (cos_m + t sin_m / sine on the phi branch, 1 otherwise)
 *   label_local = label - class_offset
 * inv_nw may be NULL: 1 / max(||w_y||, 1e-12) is then computed here from the label's weight row (so the call
 * does not have to wait for the weight normalisation).
 * Rows whose label lives on another rank get z_label = 0, dphi = 0, label_local = -1.
 * A label outside [0, C_total) sets *bad_label_flag (device int32) to 1. */
int32_t arcface_b200_label_margin(const float* x, const float* w, const float* inv_nx, const float* inv_nw,
                                  const int64_t* label, int32_t B, int32_t D, int64_t C_local,
                                  int64_t class_offset, int64_t C_total, float s, float cos_m, float sin_m,
                                  float th, float mm, int32_t easy_margin, float* t_label, float* z_label,
                                  float* dphi, int32_t* label_local, int32_t* bad_label_flag, void* stream);

/* Number of per-row partial slots arcface_b200_forward_stats writes for this shape. */
int32_t arcface_b200_forward_parts(int32_t B, int32_t D, int64_t C_local, int32_t* n_parts);

/* K2 -- cosine-logit GEMM (tcgen05 / TMEM, TMA-fed) with the margin / scale / online-softmax epilogue.
 * Replaces F.linear (arcface.py:47), the blend + scale (:58-61), CrossEntropyLoss' log-softmax and
 * torch.argmax, without writing the B x C logits.  Streams z[b, c] = s * cos[b, c] over every column
 * EXCEPT the row's label column (label_local[b], -1 = none on this shard): the label logit is computed
 * in fp32 by arcface_b200_label_margin and merged exactly by arcface_b200_finalize_rows, which keeps
 * 1 - p_label free of cancellation.  Pass label_local = NULL for the eval path (plain scaled cosines,
 * arcface.py:65-67).  Writes n_parts x B partial rows: running max, sum exp(z - max), local index of
 * the first max. */
int32_t arcface_b200_forward_stats(const uint16_t* xhat, const uint16_t* what,
                                   const int32_t* label_local, int32_t B, int32_t D, int64_t C_local, float s,
                                   float* part_max, float* part_sum, int32_t* part_arg, int32_t n_parts,
                                   void* stream);

/* K1 (class weights) + K2 in ONE launch: helper warps of the forward kernel normalise and cast the fp32 class
 * weights (outputs what / inv_nw, bit-identical to arcface_b200_normalize_cast) while the tcgen05 pipeline of the
 * same kernel consumes the rows already published, reading them back from L2; the fp32 weights cross HBM once.
 * Same partial-row outputs as arcface_b200_forward_stats (n_parts from arcface_b200_forward_parts).  Shapes the
 * fused kernel does not cover (D > 512) run the two steps as separate launches behind the same call.
 * workspace: arcface_b200_forward_fused_workspace_bytes() bytes of device memory (per-block ready counters). */
int32_t arcface_b200_forward_fused_workspace_bytes(int32_t B, int32_t D, int64_t C_local, size_t* bytes);
int32_t arcface_b200_forward_stats_fused(const uint16_t* xhat, const float* w, const int32_t* label_local, int32_t B,
                                         int32_t D, int64_t C_local, float s, uint16_t* what, float* inv_nw,
                                         float* part_max, float* part_sum, int32_t* part_arg, int32_t n_parts,
                                         void* workspace, size_t workspace_bytes, void* stream);

/* Merge partial rows of one shard (ascending class order, first max wins) into row_max / row_sum /
 * row_arg (global class id = local + class_offset). */
int32_t arcface_b200_combine_partials(const float* part_max, const float* part_sum, const int32_t* part_arg,
                                      int32_t n_parts, int32_t B, int64_t class_offset, float* row_max,
                                      float* row_sum, int64_t* row_arg, void* stream);

/* Merge the per-rank rows of the non-label columns ([n_ranks][B], rank-major; n_ranks = 1 on a single
 * GPU) with the label logit into the softmax statistics and the mean cross-entropy:
 *   z_label_out[b] = sum over ranks of rows_z_label (only the owner rank is non-zero)
 *   lse[b] = log(sum_c exp z[b, c]),  argmax[b] (lowest class id on ties; label = global class ids),
 *   one_minus_p[b] = 1 - softmax(z)[b, label]  (= S_rest / (S_rest + e^{z_label}), no cancellation),
 *   loss = mean_b(lse - z_label). */
int32_t arcface_b200_finalize_rows(const float* rows_max, const float* rows_sum, const int64_t* rows_arg,
                                   const float* rows_z_label, const int64_t* label, int32_t n_ranks, int32_t B,
                                   float* lse, int64_t* argmax, float* z_label_out, float* one_minus_p,
                                   float* loss, void* stream);

/* The same with the per-rank rows at a stride: rank r's values start rank_stride_f32 floats (rows_max, rows_sum,
 * rows_z_label) / rank_stride_i64 int64s (rows_arg) after rank r - 1's.  Lets the class-sharded head all-gather ONE
 * packed buffer per rank ([arg int64 x B | max | sum | z_label fp32 x B]: strides 5 B floats and 5 B / 2 int64s, B even)
 * and merge it in place, without unpacking kernels. */
int32_t arcface_b200_finalize_rows_strided(const float* rows_max, const float* rows_sum, const int64_t* rows_arg,
                                           const float* rows_z_label, const int64_t* label, int32_t n_ranks, int32_t B,
                                           int64_t rank_stride_f32, int64_t rank_stride_i64, float* lse,
                                           int64_t* argmax, float* z_label_out, float* one_minus_p, float* loss,
                                           void* stream);

/* Materialise out[b, c] = scale * cos[b, c] (label column overridden by z_label when given).  Eval path
 * (forward_test, scale = 1) and the debug / small-C path behind the lazy logits object. */
int32_t arcface_b200_logits(const uint16_t* xhat, const uint16_t* what, const float* z_label,
                            const int32_t* label_local, int32_t B, int32_t D, int64_t C_local, float scale,
                            float* out, int64_t ld_out, void* stream);

/* K4 -- fused cosine top-k (eval / retrieval): out_val[b, j], out_idx[b, j] = the k largest scale * cos[b, c] of row b
 * over the C rows of `what`, descending, ties by ascending index; out_idx = local index + class_offset.  Replaces
 * top-k / argmax over ArcMarginProduct.forward_test's B x C output (arcface.py:65-67) and the brute-force
 * faiss.IndexFlat(METRIC_INNER_PRODUCT).search(x, k) over L2-normalised embeddings (daodian_infer.py:225-230,
 * 295-302) without building the B x C matrix: two passes of the cosine GEMM (segment maxima -> per-row threshold ->
 * candidates above it) and a shared-memory sort.  Exact.  1 <= k <= 128, any D (multiple of 8); rows with fewer than
 * k classes are padded with (-inf, -1).  xhat / what: bf16 rows already L2-normalised (arcface_b200_normalize_cast). */
int32_t arcface_b200_topk_workspace_bytes(int32_t B, int32_t D, int64_t C, int32_t k, size_t* bytes);
int32_t arcface_b200_cosine_topk(const uint16_t* xhat, const uint16_t* what, int32_t B, int32_t D, int64_t C, int32_t k,
                                 float scale, int64_t class_offset, float* out_val, int64_t* out_idx, void* workspace,
                                 size_t workspace_bytes, void* stream);
/* Merge n_per_row (value, global index) candidates per row -- e.g. the all-gathered per-rank top-k lists of a
 * class-sharded catalogue, [B][n_per_row], n_per_row <= 16384 -- into the k best, same order as above. */
int32_t arcface_b200_topk_merge(const float* val, const int64_t* idx, int32_t B, int32_t n_per_row, int32_t k,
                                float* out_val, int64_t* out_idx, void* stream);

/* Workspace (bytes) arcface_b200_backward needs. */
int32_t arcface_b200_backward_workspace_bytes(int32_t B, int32_t D, int64_t C_local, size_t* bytes);
/* How arcface_b200_backward walks the classes: classes per scratch chunk and number of chunks
 * (each chunk is three kernel launches: dC^T producer, dW GEMM, dX GEMM). */
int32_t arcface_b200_backward_plan(int32_t B, int32_t D, int64_t C_local, int64_t* chunk_classes, int32_t* n_chunks);
/* Kernels arcface_b200_backward launches for this shape: 1 when the single-launch backward applies (B, D <= 512:
 * the dC^T / dW / dX contractions run as roles of one persistent kernel and exchange dC^T through an L2-resident
 * ring), otherwise 3 per scratch chunk. */
int32_t arcface_b200_backward_launches(int32_t B, int32_t D, int64_t C_local, int32_t* n_kernels);

/* K3 -- backward of the head + cross-entropy (loss.backward() through arcface.py:45-63).
 * Recomputes p = exp(z - lse) tile by tile from the saved row statistics (the B x C matrix is never stored;
 * with B, D <= 512 not even a full dC^T scratch: see arcface_b200_backward_launches), forms
 *   dC[b, c] = s * grad_scale * p            (c != label)
 *   dC[b, y] = -s * grad_scale * one_minus_p[b] * dphi[b]      (one_minus_p from arcface_b200_finalize_rows)
 * and runs dXhat = dC . What (accumulated into the zeroed dxhat) and dWhat = dC^T . Xhat; the epilogue of
 * the dW GEMM applies the normalise backward  dW[c] = (dWhat[c] - (what[c] . dWhat[c]) what[c]) * inv_nw[c].
 * grad_scale = upstream grad of the mean loss / global batch size; when grad_loss_dev (nullable, DEVICE
 * float scalar) is given, the effective scale is grad_scale * *grad_loss_dev, so an autograd caller never
 * has to synchronise to read the upstream gradient.  dxhat is this rank's partial (sum over its classes)
 * and is reduce-scattered by the caller when the head is class-sharded. */
int32_t arcface_b200_backward(const uint16_t* xhat, const uint16_t* xhat_t, int64_t ld_t, const uint16_t* what,
                              const float* inv_nw, const float* lse, const float* one_minus_p, const float* dphi,
                              const int32_t* label_local, int32_t B, int32_t D, int64_t C_local, float s,
                              float grad_scale, const float* grad_loss_dev, float* dxhat, float* dw,
                              void* workspace, size_t workspace_bytes, void* stream);

/* arcface_b200_backward with a precision mode.  prec = ARCFACE_B200_PREC_BF16X3: `xhat` [B][3 D] and `what`
 * [C_local][3 D] are the rows written by arcface_b200_normalize_cast3 (orders 0 and 1), `xhat_t` its transposed output
 * [D][3 ld_t] = [hi^T | lo^T | hi^T] with ld_t = B rounded up to 64 and ZERO padding columns.  The probabilities are
 * recomputed from the three-term product; dC leaves its epilogue as a bf16 pair hi + lo ([hi | hi | lo] blocks of the
 * scratch), the dW GEMM contracts over the three blocks, dX runs as three launches (hi.W_hi + hi.W_lo + lo.W_hi) and a
 * row pass adds the lo part of the weight projection: gradients within ~1e-5 relative of fp32 (B <= 1024; larger
 * batches keep bf16 dC and the hi operands). */
int32_t arcface_b200_backward_prec(const uint16_t* xhat, const uint16_t* xhat_t, int64_t ld_t, const uint16_t* what,
                                   const float* inv_nw, const float* lse, const float* one_minus_p, const float* dphi,
                                   const int32_t* label_local, int32_t B, int32_t D, int64_t C_local, float s,
                                   float grad_scale, const float* grad_loss_dev, float* dxhat, float* dw,
                                   void* workspace, size_t workspace_bytes, int32_t prec, void* stream);

/* The step before the head in the two-stream model (multimodal_classifier.py:50-56):
 *   out = cat(F.normalize(a), F.normalize(b), dim = 1),  a [B][D1], b [B][D2], out [B][D1 + D2] fp32
 * in one pass; inv1 / inv2 [B] receive 1 / max(||.||, 1e-12).  D1, D2 multiples of 4. */
int32_t arcface_b200_two_stream_concat(const float* a, const float* b, int32_t B, int32_t D1, int32_t D2, float* out,
                                       float* inv1, float* inv2, void* stream);
/* Its backward: da = (g_a - e_a (e_a . g_a)) * inv1, db likewise, with e = the halves of `emb` (the forward's output)
 * and g = the halves of `grad` [B][D1 + D2]. */
int32_t arcface_b200_two_stream_concat_bwd(const float* emb, const float* inv1, const float* inv2, const float* grad,
                                           int32_t B, int32_t D1, int32_t D2, float* da, float* db, void* stream);

/* Normalise backward for the embeddings: dx[b] = (dxhat[b] - (xhat[b] . dxhat[b]) xhat[b]) * inv_nx[b]
 * with xhat = x * inv_nx in fp32. */
int32_t arcface_b200_normalize_bwd_x(const float* x, const float* inv_nx, const float* dxhat, int32_t B,
                                     int32_t D, float* dx, void* stream);

/* In-place a[0..na) *= *scale_dev, b[0..nb) *= *scale_dev (na, nb multiples of 4; either may be 0).  Used when the
 * backward was run ahead of time with an upstream gradient of 1 (CUDA-graph replay of forward + backward in one
 * launch): the gradients are linear in the upstream gradient of the scalar loss, so loss.backward() only has to
 * apply it.  When *scale_dev == 1.0f -- loss.backward() on the head's own loss -- the kernel returns immediately. */
int32_t arcface_b200_scale_grads(float* a, int64_t na, float* b, int64_t nb, const float* scale_dev, void* stream);

/* dst[0..n) = src[0..n) * *scale_dev (always written), b[0..nb) *= *scale_dev unless the factor is exactly 1
 * (n, nb multiples of 4).  The graph-replay path hands `dst` (a fresh dX) to autograd and keeps `src` as its static
 * buffer: one launch instead of arcface_b200_scale_grads + a device copy. */
int32_t arcface_b200_scale_copy(const float* src, float* dst, int64_t n, float* b, int64_t nb, const float* scale_dev,
                                void* stream);

/* dst = [x (b x D fp32) | y (b int64)]: the step's inputs into the packed static buffer of a captured graph (the
 * layout the class-sharded head's gather sends as is) in one launch. */
int32_t arcface_b200_pack_xy(const float* x, const int64_t* y, int32_t b, int32_t D, void* dst, void* stream);

/* Fused head optimiser: one torch.optim.AdamW step (decoupled weight decay, bias correction for step number `step`
 * >= 1) on the fp32 class-weight rows, in place on w / exp_avg / exp_avg_sq, and -- when what / inv_nw are given --
 * the NEXT forward's K1 in the same pass: what = bf16(w_new / max(||w_new||, 1e-12)), inv_nw = 1 / max(||w_new||, 1e-12).
 * Replaces AdamW(model.classifier.parameters()).step() (nlp_classifier_train.py:94-97, 131-133) plus
 * F.normalize(self.weight) of the following forward (arcface.py:47): 30 bytes per weight instead of 28 + 6. */
int32_t arcface_b200_adamw_normalize(float* w, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t rows,
                                     int32_t D, double lr, double beta1, double beta2, double eps, double weight_decay,
                                     int64_t step, uint16_t* what, float* inv_nw, void* stream);

/* One-shot exchange over peer-mapped memory (NVLink): stores bytes_per_peer bytes from src (+ r * src_stride for peer r;
 * src_stride = 0 sends the same bytes to everyone = all-gather, src_stride = bytes_per_peer = scatter) into slot `rank`
 * of every rank's receive buffer (peer_bufs[r] + rank * slot_stride), raises this call's flag on every rank and
 * returns (stream order) once every rank's flag for the same call has arrived here.  peer_bufs / peer_flags: HOST
 * arrays of `world` device addresses valid in this process (torch.distributed._symmetric_memory buffer_ptrs);
 * flags = uint32 [8][16] per rank, zeroed once; sync_dev = uint32 [16] of local device memory, zeroed once.  All ranks
 * must issue the same sequence of calls per channel.  error_word (nullable; device or mapped pinned HOST memory): set
 * to 1 + channel when a peer's flag has not arrived after a few seconds of polling -- the kernel then finishes without
 * it (the step's results are meaningless) instead of hanging the GPU or trapping the context; the caller examines the
 * word before its next exchange.  Replaces the latency-bound NCCL all-gather / reduce-scatter of
 * the class-sharded step (sharded.py; nn.DataParallel's gather / reduce in the reference). */
int32_t arcface_b200_p2p_exchange(const void* src, size_t bytes_per_peer, size_t src_stride, const uint64_t* peer_bufs,
                                  const uint64_t* peer_flags, int32_t rank, int32_t world, size_t slot_stride,
                                  int32_t channel, uint32_t* sync_dev, uint32_t* error_word, void* stream);

/* The all-gather form of arcface_b200_p2p_exchange (src_stride = 0) with a split receive layout: the first `split` bytes
 * of every rank's message land contiguously in rank order at the start of each receive buffer ([world][split]), the
 * remaining bytes contiguously behind them ([world][bytes_per_peer - split]).  The class-sharded head gathers the
 * byte-packed (x | labels) of every rank this way and reads one [B][D] matrix and one [B] label vector straight out
 * of the receive buffer.  split = 0: plain slots, as arcface_b200_p2p_exchange. */
int32_t arcface_b200_p2p_gather_split(const void* src, size_t bytes_per_peer, size_t src_stride, size_t split,
                                      const uint64_t* peer_bufs, const uint64_t* peer_flags, int32_t rank,
                                      int32_t world, size_t slot_stride, int32_t channel, uint32_t* sync_dev, uint32_t* error_word,
                                      void* stream);
/* normalize_bwd_x over the SUM of n_parts partial dxhat buffers ([n_parts][B][D], part_stride floats apart), summed in
 * index order: the reduce half of the reduce-scatter, fused. */
int32_t arcface_b200_normalize_bwd_x_sum(const float* x, const float* inv_nx, const float* parts, int32_t n_parts,
                                         int64_t part_stride, int32_t B, int32_t D, float* dx, void* stream);

/* One-call step for hosts without torch: HOST embeddings / labels in, HOST loss / argmax / dx out; the
 * class weights and their gradient stay resident on the device (w, dw are DEVICE pointers, fp32 [C x D]).
 * Copies in and out are part of the call; it returns after the results have landed in the host buffers.
 * device_ws must hold arcface_b200_step_workspace_bytes() bytes. */
int32_t arcface_b200_step_workspace_bytes(int32_t B, int32_t D, int64_t C, size_t* bytes);
int32_t arcface_b200_step_host(const float* x_host, const int64_t* label_host, const float* w_dev, int32_t B,
                               int32_t D, int64_t C, float s, float m, int32_t easy_margin, float grad_loss,
                               float* loss_host, int64_t* argmax_host, float* dx_host, float* dw_dev,
                               void* device_ws, size_t device_ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ARCFACE_B200_H_ */
