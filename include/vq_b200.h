/*
 * vq_b200.h -- C ABI of libvq_b200.so, the B200 (sm_100a) implementation of the VQ codebook
 * quantiser hot path of pranoyr/attention-models.
 *
 * The reference has no FFI / plugin layer for this path: its boundary is the Python nn.Module
 * surface of the two `Codebook` classes.  Each entry point below names the
 * reference code it replaces (paths relative to /root/reference).  The Python drop-in modules in
 * attention-models_b200/vq_b200/ bind these symbols with ctypes (INTEGRATION.md shows the stub a
 * maintainer of the reference would add).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless its name ends in _host
 *   - all arrays are caller-allocated; the library owns no device memory (vq_peer_alloc allocates the
 *     IPC-exportable exchange buffer on the caller's explicit request; the caller frees it)
 *   - every call is asynchronous on `stream` (a cudaStream_t) and re-entrant per stream and per device; the two
 *     exceptions keep per-device state inside the library and are single-caller: vq_host_step (its copy streams and
 *     events) and the vq_profile_* measurement hooks (process-global event lists)
 *   - return value: 0 = OK, non-zero = error; vq_last_error() gives the message (thread-local)
 *   - nothing throws across the ABI; there is NO CPU fallback: without a CUDA device every compute
 *     call returns VQ_ERR_CUDA
 *   - floats are IEEE fp32; indices are int64 (torch.argmin's dtype)
 *   - token-major layout: (T, D) row-major.  NCHW layout: (b, D, hw) with T = b*hw and the flat
 *     token order (b, h, w) the reference uses (models/vqgan.py:153)
 */
#ifndef VQ_B200_H
#define VQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VQ_ABI_VERSION 6

#if defined(__GNUC__)
#define VQ_API __attribute__((visibility("default")))
#else
#define VQ_API
#endif

/* which Codebook: decides where beta sits in the loss and whether decode renormalises */
#define VQ_FORM_VIT   0   /* models/vitvqgan.py:140-176 */
#define VQ_FORM_VQGAN 1   /* models/vqgan.py:138-182   */
#define VQ_FORM_VQGAN_L2 2 /* the CNN form WITHOUT l2 normalisation: plain squared-L2 nearest code on the raw vectors,
                              d = (|z|^2 + |e_k|^2) - 2 z.e_k, q = E[idx], loss / STE / layouts as VQ_FORM_VQGAN.  Not a form
                              of the reference (both its Codebooks normalise); BASELINE.json's north_star names it.  Exhaustive
                              fp32 search only (the tensor-core filters' error bounds assume unit rows); single GPU or the
                              collective exchange.  `cb` must have been prepared by this form (vq_forward(weight != NULL)
                              or vq_codebook_prepare_raw).                                                      */

#define VQ_LAYOUT_TOKEN_MAJOR 0   /* (T, D)      -- ViT form input/output            */
#define VQ_LAYOUT_NCHW        1   /* (b, D, h*w) -- VQGAN form input/output          */

/* flags for vq_forward */
#define VQ_FLAG_INDICES_ONLY 1    /* encode_imgs fast path: only idx (and hist) are produced   */
#define VQ_FLAG_EXACT_SCAN   2    /* force the exhaustive fp32 SIMT search (no tensor cores)   */
#define VQ_FLAG_KEEP_STATS   4    /* hist, stats and seg_sums accumulate (the caller zeroed them once):
                                     one batch can be streamed through in token chunks           */
#define VQ_FLAG_IDX32        8    /* with VQ_FLAG_INDICES_ONLY: `idx` is T int32 ...                            */
#define VQ_FLAG_IDX16        16   /* ... or T uint16 (codebook_size <= 65536): the token wire format for the
                                     consumers below (vq_gather_tokens, vq_token_embed_tokens), 2 or 4 instead
                                     of 8 bytes per token on every copy behind encode_imgs                       */

/* error codes */
#define VQ_OK            0
#define VQ_ERR_ARG       1
#define VQ_ERR_WORKSPACE 2
#define VQ_ERR_CUDA      3
#define VQ_ERR_INDEX     4        /* index >= K handed to vq_gather (reference: IndexError)     */

/* slots of the int64 `stats` array written by vq_forward (device memory, VQ_STATS_LEN entries) */
#define VQ_STAT_NEAR_TIE_ROWS   0 /* rows whose two best fp32 distances differ by < 1e-6 relative */
#define VQ_STAT_AMBIGUOUS_ROWS  1 /* rows the tensor-core pass could not decide alone (rescored over >1 cell) */
#define VQ_STAT_FALLBACK_ROWS   2 /* rows sent to the exhaustive fp32 search (a row the split filter
                                     leaves undecided in two code ranges is counted twice)        */
#define VQ_STAT_LOSS_FIXED      3 /* sum over tokens of sum_j (q - zn)^2, fixed point 2^-24      */
#define VQ_STAT_BAD_INDEX       4 /* vq_gather: count of out-of-range indices                    */
#define VQ_STAT_NONFINITE       5 /* non-finite loss partials (NaN/Inf rows): the loss is NaN     */
#define VQ_STAT_PEER_TIMEOUT    6 /* sharded backward gave up: a peer never published its step within the time-out, or
                                     the caller's slot / epoch is out of step with the device's count; grad_weight
                                     and loss of that call are NaN (see vq_peer_configure)            */
#define VQ_STATS_LEN            8

VQ_API int         vq_abi_version(void);
VQ_API const char* vq_last_error(void);

/* 1 if vq_forward would run the tcgen05 tensor-core search for this shape, 0 if the exhaustive fp32
 * SIMT search (small or unsupported shapes).  Pure host logic.                                     */
VQ_API int vq_uses_tensor_cores(int64_t T, int K, int D);

/* Number of SMs / compute capability of the current device (0 on failure). */
VQ_API int vq_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- codebook preparation ------------------------------------------------------------------
 * Replaces `embedd_norm = l2_norm(self.embedding.weight)` and `torch.sum(embedd_norm**2, dim=1)`
 * (models/vitvqgan.py:154,158; models/vqgan.py:155,158).  Produces, inside `cb` (opaque, size from
 * vq_codebook_bytes): the unit codes in fp32, their squared norms, max(||E_k||, eps), and the fp16
 * copy the tensor-core search streams through TMA.  Re-run whenever the weights change; a frozen
 * tokeniser prepares once.                                                                        */
VQ_API int vq_codebook_bytes(int K, int D, size_t* out);
VQ_API int vq_codebook_prepare(const float* weight, int K, int D, void* cb, size_t cb_bytes, void* stream);
/* the same blob for VQ_FORM_VQGAN_L2: the codes as they are, sum(E_k^2), denominators of 1 */
VQ_API int vq_codebook_prepare_raw(const float* weight, int K, int D, void* cb, size_t cb_bytes, void* stream);

/* ---- forward -------------------------------------------------------------------------------
 * Replaces Codebook.forward (models/vitvqgan.py:151-171, models/vqgan.py:148-176):
 *   zn = l2norm(z); idx = argmin_k ((|zn|^2 + |en_k|^2) - 2 zn.en_k); q = l2norm(E[idx]);
 *   loss; z_q = zn + (q - zn).
 * weight      : NULL if `cb` is already prepared (a frozen tokeniser prepares once); else the K*D raw weights, and
 *               this call prepares `cb` itself (a training step: the weights changed) -- in the same launch as the
 *               token rows where the shape allows (token-major, D < 128), so the step has no separate prepare.
 * z, z_q      : `layout`, fp32.  z_q may be NULL with VQ_FLAG_INDICES_ONLY.
 * idx         : T int64, flat token order.
 * loss        : 1 float, the reference's loss with the mean taken over `n_elem_total` elements
 *               (pass T*D on one GPU; a token-sharded job passes the global count and finishes the
 *               loss after its all-reduce with vq_loss_finalize).  May be NULL.
 * hist        : K int32 code-usage counts of THIS call (overwritten), or NULL.
 * stats       : VQ_STATS_LEN int64 (overwritten), or NULL.
 * saved_zn    : T*D floats, token-major unit rows kept for vq_backward, or NULL.
 * saved_denom : T floats max(||z_t||, eps) kept for vq_backward, or NULL.
 * seg_sums    : K*D + K int64, or NULL.  The codebook-gradient segment sums S_k = sum_{t: idx_t = k} (q_k - zn_t)
 *               at fixed-point scale 2^30, followed by K per-code counts of non-finite terms -- what
 *               vq_backward_codebook consumes.  They depend on neither the upstream gradient nor g_loss, so a
 *               training step accumulates them here, in the pass that already holds q - zn (64-bit integer
 *               reductions: exact, order-free, deterministic), and a token-sharded job can start their exchange
 *               before the backward.  Overwritten unless VQ_FLAG_KEEP_STATS.
 * ws          : scratch of at least vq_workspace_bytes(T, K, D, flags).                            */
VQ_API int vq_workspace_bytes(int64_t T, int K, int D, int flags, size_t* out);
VQ_API int vq_forward(const float* z, int layout, int64_t T, int64_t hw,
               const float* weight, void* cb, int K, int D, int form, float beta, int flags, int64_t n_elem_total,
               float* z_q, void* idx, float* loss, int32_t* hist, int64_t* stats,
               float* saved_zn, float* saved_denom, int64_t* seg_sums,
               void* ws, size_t ws_bytes, void* stream);

/* ---- pre_quant fused into the forward ------------------------------------
 * Replaces `enc = self.pre_quant(enc); self.codebook(enc)` of ViTVQGAN.forward / encode_imgs
 * (models/vitvqgan.py:185, 192-193, 207-208; pre_quant = nn.Linear(dim, codebook_dim)):
 *   z = x W_pre^T + b_pre   (x: (T, C) encoder rows, W_pre: (D, C) = pre_quant.weight, b_pre: D floats or NULL)
 * is formed inside the token preparation -- x is read once, z never travels through HBM -- and everything behind it is
 * vq_forward's (same arguments, same outputs, form = VQ_FORM_VIT, token-major).  The GEMM runs on the tensor cores in the
 * 3xTF32 split (fp32-GEMM-level rounding, in an order of its own like any two fp32 GEMMs): z agrees with F.linear to a
 * few fp32 ulps of sum|x_c W_dc|; an index can differ from F.linear + vq_forward only on a row whose two best codes are
 * closer than that rounding (tests count them like near-ties).  Given the same z all outputs are bit-identical to
 * vq_forward's.
 * z_out: optional (T, D) copy of z (NULL: not written -- the backward needs saved_zn / saved_denom only).
 * Supported: D = 32, C a multiple of 64 up to 768 (vq_prequant_supported; else VQ_ERR_ARG: project with a GEMM and call
 * vq_forward).  x must be 16-byte aligned.  The backward is vq_backward for grad_z followed by the caller's two GEMMs
 * (grad_x = grad_z W_pre, grad_W_pre = grad_z^T x).                                                              */
VQ_API int vq_prequant_supported(int C, int D);
VQ_API int vq_forward_projected(const float* x, int C, const float* w_pre, const float* b_pre, int64_t T,
               const float* weight, void* cb, int K, int D, int form, float beta, int flags, int64_t n_elem_total,
               float* z_q, void* idx, float* loss, int32_t* hist, int64_t* stats,
               float* saved_zn, float* saved_denom, int64_t* seg_sums, float* z_out,
               void* ws, size_t ws_bytes, void* stream);

/* loss = the reference's two-term expression from the fixed-point sum (after an all-reduce). */
VQ_API int vq_loss_finalize(const int64_t* loss_fixed, int64_t n_elem_total, int form, float beta,
                     float* loss, void* stream);

/* ---- backward ------------------------------------------------------------------------------
 * Replaces what autograd derives from Codebook.forward:
 *   g_zn = G + g_loss*c1*2(zn - q)/N ; grad_z = NB(z, g_zn)
 *   S_k  = sum_{t: idx_t = k} (q_k - zn_t)          (deterministic segmented sum, fixed point)
 *   grad_E[k] = NB(E_k, g_loss*c2*(2/N) S_k)        c1,c2 = (beta,1) ViT / (1,beta) VQGAN
 * vq_backward_tokens writes grad_z (layout of z) and, if asked (seg_sums != NULL; not needed when
 * vq_forward already produced them), the int64 fixed-point segment sums from idx alone: tokens are
 * bucketed by code and every bucket summed in a fixed order (K*D sums at scale 2^30 followed by K
 * per-code counts of non-finite contributions; K*D + K int64 in all, overwritten; bit-identical to
 * vq_forward's).  A token-sharded job adds the seg_sums of all ranks (integer sum: exact, order-free)
 * before vq_backward_codebook; one GPU calls them back to back.
 * g_zq may be NULL (no upstream gradient through z_q); grad_z / seg_sums may be NULL (not wanted).
 * hist: the K int32 code-usage counts vq_forward wrote for the same idx (saves a recount), or NULL.
 * g_loss: DEVICE pointer to d(objective)/d(loss) (what autograd hands over), NULL means 1.0.
 * vq_backward_codebook: with `loss` != NULL the same launch also writes the loss from `stats` (what
 * vq_loss_finalize computes; for callers that did not ask vq_forward for it); both may be NULL.     */
VQ_API int vq_backward_workspace_bytes(int64_t T, int K, int D, size_t* out);
VQ_API int vq_backward_tokens(const float* g_zq, int layout, int64_t T, int64_t hw,
                       const float* saved_zn, const float* saved_denom, const int64_t* idx, const int32_t* hist,
                       const void* cb, int K, int D, int form, float beta, const float* g_loss,
                       int64_t n_elem_total,
                       float* grad_z, int64_t* seg_sums,
                       void* ws, size_t ws_bytes, void* stream);
VQ_API int vq_backward_codebook(const int64_t* seg_sums, const void* cb, int K, int D, int form, float beta,
                         const float* g_loss, int64_t n_elem_total, float* grad_weight,
                         const int64_t* stats, float* loss, void* stream);

/* The backward of a step whose seg_sums came from vq_forward, in ONE launch where the layout allows (token-major):
 * grad_z as vq_backward_tokens, grad_weight (and, with `loss` != NULL, the loss from `stats`) as
 * vq_backward_codebook.  grad_z may be NULL.  ws as for vq_backward_tokens (only used for the NCHW layout).        */
VQ_API int vq_backward(const float* g_zq, int layout, int64_t T, int64_t hw,
                const float* saved_zn, const float* saved_denom, const int64_t* idx,
                const void* cb, int K, int D, int form, float beta, const float* g_loss, int64_t n_elem_total,
                const int64_t* seg_sums, const int64_t* stats,
                float* grad_z, float* grad_weight, float* loss,
                void* ws, size_t ws_bytes, void* stream);

/* ---- token-sharded job: the backward's one exchange, fused with the codebook gradient -------------
 * Replaces DDP's all-reduce of codebook.embedding.weight.grad (trainers/vitgqgan.py:184,
 * trainers/utils/base_trainer.py:29-33) for one process per GPU on one NVSwitch box.  Every rank owns an
 * exchange buffer of vq_exchange_bytes(K, D) bytes that all peers map (CUDA IPC):
 *   vq_peer_alloc   cudaMalloc + zero + export: *dev_ptr and a 64-byte IPC handle to send to the peers
 *   vq_peer_open    map a peer's handle (enables NVLink peer access); vq_peer_close unmaps it
 *   vq_peer_free    free the own buffer (after every peer closed it)
 * A buffer holds two slots (step parity); vq_exchange_slot gives, for `slot`, the device pointers to pass to
 * vq_forward as seg_sums / stats / hist so that the forward writes its partials straight into the buffer.
 * vq_backward_codebook_sharded then runs ONE kernel per rank that publishes "step `epoch` ready" to the
 * peers, waits for theirs, reads every rank's integer partials over NVLink, adds them (exact, order-free:
 * all ranks get bit-identical totals) and writes grad_weight (K*D), the global histogram (K int64, or
 * NULL), the global loss (or NULL) and the summed stats (VQ_STATS_LEN int64, or NULL).
 * peer_bufs: HOST array of `world` device pointers, peer_bufs[rank] the own buffer; world <=
 * VQ_PEER_MAX_RANKS.  The step number lives in the exchange buffer (the kernel bumps it), so a launch carries no
 * per-step host state and can be captured in a CUDA graph: `slot` must alternate 0, 1, 0, ... from the first
 * call on; `epoch` is the caller's own count of the calls (1, 2, ...) for a consistency check, or 0 to skip it
 * (graph capture).  A mismatch is reported like a peer time-out.  n_elem_total is the GLOBAL element count.
 * world = 1 needs no peers.                                                                              */
#define VQ_PEER_MAX_RANKS 16
#define VQ_IPC_HANDLE_BYTES 64
VQ_API int vq_exchange_bytes(int K, int D, size_t* out);
VQ_API int vq_exchange_slot(void* exchange_buf, int K, int D, int slot, int64_t** seg_sums, int64_t** stats,
                            int32_t** hist);
VQ_API int vq_peer_alloc(size_t bytes, void** dev_ptr, void* ipc_handle_out);
VQ_API int vq_peer_open(const void* ipc_handle, void** dev_ptr);
VQ_API int vq_peer_close(void* dev_ptr);
VQ_API int vq_peer_free(void* dev_ptr);
/* Time-out and failure reporting of the exchange.  A rank waits at most `timeout_ms` (0 = the default, 10 minutes --
 * the order of a collective library's watchdog) for a peer to publish its step.  Giving up is FATAL for the exchange,
 * never silent: the launch writes NaN into grad_weight and loss, bumps stats_total[VQ_STAT_PEER_TIMEOUT], does not
 * advance the device-side step count (every later launch on this buffer fails the same way) and stores 1 into
 * `*abort_flag` -- an int in pinned, device-visible HOST memory that the caller polls without synchronising (NULL: none).
 * A launch whose slot / epoch disagrees with the device-side count is treated the same way.  Recovery is collective:
 * with no exchange kernel in flight on any rank, every rank calls vq_peer_resync on its own buffer between two
 * barriers of the job (flags, counters and the step count return to zero; the next call is step 1, slot 0).     */
VQ_API int vq_peer_configure(void* own_exchange_buf, int64_t timeout_ms, void* abort_flag, void* stream);
VQ_API int vq_peer_resync(void* own_exchange_buf, void* stream);
VQ_API int vq_backward_codebook_sharded(const void* const* peer_bufs, int world, int rank, int slot, uint32_t epoch,
                                        const void* cb, int K, int D, int form, float beta, const float* g_loss,
                                        int64_t n_elem_total, float* grad_weight, int64_t* hist_total, float* loss,
                                        int64_t* stats_total, void* stream);

/* The whole backward of a token-sharded step (token-major rows) in ONE launch: the first blocks run the exchange +
 * codebook gradient above, the others grad_z as vq_backward_tokens -- no second stream, no cross-stream events; the
 * token backward fills the SMs the exchange leaves idle while it waits for the other GPUs.                        */
VQ_API int vq_backward_sharded(const void* const* peer_bufs, int world, int rank, int slot, uint32_t epoch,
                               const float* g_zq, int64_t T, const float* saved_zn, const float* saved_denom,
                               const int64_t* idx, const void* cb, int K, int D, int form, float beta, const float* g_loss,
                               int64_t n_elem_total, float* grad_z, float* grad_weight, int64_t* hist_total, float* loss,
                               int64_t* stats_total, void* stream);

/* ---- decode --------------------------------------------------------------------------------
 * Replaces Codebook.indices_to_embeddings: models/vitvqgan.py:173-176 (normalise=1, token-major)
 * and models/vqgan.py:178-182 (normalise=0, NCHW out with hw = n).  normalise=1 reads the unit
 * codes from `cb`; normalise=0 reads `weight`.  Out-of-range indices are counted in
 * stats[VQ_STAT_BAD_INDEX] (the Python side raises IndexError like the reference) and produce 0. */
VQ_API int vq_gather(const int64_t* idx, int64_t T, int64_t hw, const float* weight, const void* cb,
              int K, int D, int normalise, int layout_out, float* out, int64_t* stats, void* stream);
/* ... from tokens in a narrow wire format: token_bits = 16 (uint16), 32 (int32) or 64 (int64) */
VQ_API int vq_gather_tokens(const void* tokens, int token_bits, int64_t T, int64_t hw, const float* weight, const void* cb,
              int K, int D, int normalise, int layout_out, float* out, int64_t* stats, void* stream);
/* tokens between the wire formats (uint16 / int32 / int64); values that do not fit the target are the caller's error */
VQ_API int vq_tokens_convert(const void* in, int in_bits, void* out, int out_bits, int64_t T, void* stream);

/* ---- post_quant fused into the decode -------------------------------------
 * Replaces `self.post_quant(self.codebook.indices_to_embeddings(indices))` of decode_indices
 * (models/vitvqgan.py:199-200 with nn.Linear(codebook_dim, dim); models/vqgan.py:241-242 with nn.Conv2d(dim, dim, 1)):
 * the projection of a code does not depend on the token, so it is applied to the K codes once,
 *   table[k] = W_post y_k + b_post     y_k = l2norm(E_k) (normalise=1, from `cb`) or E_k (normalise=0, from `weight`)
 * (W_post: (C, D) row-major = post_quant.weight, also the (C, D, 1, 1) conv kernel; b_post: C floats or NULL; fp32 fma
 * chain over d then the bias), and decode_indices becomes ONE gather of (T, C) rows -- token-major, or (b, C, hw) for
 * the CNN form -- from a table that stays in L2 (16 MB at K = 8192, C = 512).  Re-run vq_project_codebook when the
 * codebook or post_quant change.  Results agree with F.linear / F.conv2d on the gathered codes to fp32 rounding of a
 * D-term dot product.  vq_gather_projected: any C > 0 (C % 4 == 0 for token-major), tokens as for vq_gather_tokens,
 * out-of-range tokens counted in stats[VQ_STAT_BAD_INDEX] and written as 0.                                      */
VQ_API int vq_project_codebook(const float* weight, const void* cb, int K, int D, int normalise,
                               const float* w_post, const float* b_post, int C, float* table, void* stream);
VQ_API int vq_gather_projected(const void* tokens, int token_bits, int64_t T, int64_t hw, const float* table, int K, int C,
                               int layout_out, float* out, int64_t* stats, void* stream);

/* ---- first consumer of the tokens ------------------------------------------
 * The mask-fill + token-embedding lookup the generative models do right behind encode_imgs, in one pass:
 *   input_ids = tokens.masked_fill(mask, mask_token_id)        models/muse.py:149, models/maskgit.py:132
 *   labels    = tokens.masked_fill(~mask, ignore_index)        models/muse.py:150, models/maskgit.py:131
 *   embeds[t] = table[input_ids[t]] + pos[t mod n_per_seq]     models/muse.py:90-91, models/maskgit.py:80-81
 * tokens: T int64; mask: T bytes (0 / non-zero) or NULL (nothing masked); table: (vocab, dim) fp32 with dim % 4 == 0;
  * pos: (n_per_seq, dim) fp32 or NULL; embeds (T, dim), input_ids, labels: any may be NULL.  The backward with respect
 * to the table is vq_embedding_backward over input_ids.  Out-of-range ids are counted in stats[VQ_STAT_BAD_INDEX] and embed to 0.                     */
VQ_API int vq_token_embed(const int64_t* tokens, const uint8_t* mask, int64_t T, int64_t n_per_seq,
                          int64_t mask_token_id, int64_t ignore_index, const float* table, int64_t vocab, int dim,
                          const float* pos, float* embeds, int64_t* input_ids, int64_t* labels, int64_t* stats,
                          void* stream);
VQ_API int vq_token_embed_tokens(const void* tokens, int token_bits, const uint8_t* mask, int64_t T, int64_t n_per_seq,
                          int64_t mask_token_id, int64_t ignore_index, const float* table, int64_t vocab, int dim,
                          const float* pos, float* embeds, int64_t* input_ids, int64_t* labels, int64_t* stats,
                          void* stream);   /* tokens: uint16 / int32 / int64 by token_bits; input_ids, labels stay int64 */

/* The autoregressive consumer (Parti, models/parti.py:98-106): the decoder input of a (b, n) token batch is
 *   embeds[b, 0] = start_token ;  embeds[b, i] = table[tokens[b, i - 1]] + pos[i - 1]   (i = 1 .. n - 1)
 * i.e. `token_emb(tokens[:, :-1])`, `+ pe[:n-1]` (models/positional_encoding.py:40-41; the dropout behind it is the
 * caller's) and `cat(start_token, ...)` in one pass; the labels are the tokens themselves.  T = b * n_per_seq.      */
VQ_API int vq_token_embed_causal(const int64_t* tokens, int64_t T, int64_t n_per_seq, const float* table, int64_t vocab,
                                 int dim, const float* pos, const float* start, float* embeds, int64_t* stats, void* stream);
VQ_API int vq_token_embed_causal_tokens(const void* tokens, int token_bits, int64_t T, int64_t n_per_seq, const float* table,
                                 int64_t vocab, int dim, const float* pos, const float* start, float* embeds, int64_t* stats,
                                 void* stream);

/* Backward of an embedding lookup, deterministic: grad_table[ids[j]] += grad_out[row(j)], accumulated as 64-bit
 * integers at a per-call fixed-point scale derived from max |grad_out| (exact, order-free sums), converted once.
 * Replaces what autograd derives for nn.Embedding behind the tokens (models/muse.py:90, models/maskgit.py:80,
 * models/parti.py:100) and for Codebook.indices_to_embeddings (models/vitvqgan.py:173-176, models/vqgan.py:178-182).
 *   row(j) = (j / ids_per_seq) * rows_per_seq + j % ids_per_seq + row_shift     (0, 0, 0: row(j) = j; Parti's shifted
 *            input: ids_per_seq = n - 1 over tokens[:, :-1] made contiguous, rows_per_seq = n, row_shift = 1)
 *   hw > 0 : grad_out is (b, dim, hw) (the VQGAN decode layout), id j = b * hw + p pairs with grad_out[b, :, p]
 *   ids outside [0, vocab) take no gradient;  grad_out holds n_rows rows of `dim` floats (dim % 4 == 0)
 *   cb != NULL (the prepared codebook, vocab = K, dim = D): the lookup was l2norm(E[i]) -- grad_table is then the
 *            codebook gradient NB(E_k, sum of the upstream rows)
 * grad_table (vocab, dim) is overwritten.  ws: vq_embedding_backward_bytes(vocab, dim) of device scratch.          */
VQ_API int vq_embedding_backward_bytes(int64_t vocab, int dim, size_t* out);
VQ_API int vq_embedding_backward(const int64_t* ids, int64_t n_ids, int64_t ids_per_seq, int64_t rows_per_seq,
                                 int64_t row_shift, const float* grad_out, int64_t n_rows, int64_t hw, int64_t vocab, int dim,
                                 const void* cb, float* grad_table, void* ws, size_t ws_bytes, void* stream);

/* ---- measurement hooks (bench.py) -------------------------------------------------------------
 * Between vq_profile_begin() and vq_profile_end() the library counts all kernel launches and brackets the kernels
 * of every `sample_every`-th step (a vq_forward and the calls that follow it) with CUDA events on the caller's
 * stream, one slot per kernel family (those of `slot_mask`).  An event pair costs about 5 us of an otherwise
 * gap-free step, which is why a timed region samples and brackets only the search kernels (bench.py), and times
 * the other families in a separate pass.
 * vq_profile_end synchronises and returns the summed time and launch count of the nearest-code search (the
 * dominant kernel: the tensor-core filter, or the exhaustive scan) and the number of kernel launches of all kinds;
 * vq_profile_slot then returns (and clears) any other slot.                                          */
#define VQ_PROFILE_PREP_CODEBOOK   0
#define VQ_PROFILE_PREP_TOKENS     1
#define VQ_PROFILE_SEARCH          2   /* the tensor-core filter (without it: the exhaustive scan)                   */
#define VQ_PROFILE_EXACT_FINISH    3   /* everything behind the filter: exact rescoring, undecided rows, finish pass  */
#define VQ_PROFILE_TAIL            4   /* overflow of the undecided-row list (normally an empty launch)            */
#define VQ_PROFILE_BACKWARD_TOKENS 5
#define VQ_PROFILE_CODEBOOK_GRAD   6   /* incl. the fused peer exchange of a token-sharded job                      */
#define VQ_PROFILE_SLOTS           7
/* Kernels this library has launched so far in this process (a launch recorded into a CUDA graph counts once, at
 * capture; its replays are the caller's to count).                                                    */
VQ_API int64_t vq_kernel_launches(void);
VQ_API int vq_profile_begin(int sample_every, unsigned slot_mask /* bit i = VQ_PROFILE_* slot i; 0 = all */);
VQ_API int vq_profile_end(double* search_ms_total, int64_t* search_launches, int64_t* kernel_launches);
VQ_API int vq_profile_slot(int slot, double* ms_total, int64_t* launches);

/* ---- host-buffer entry points (end-to-end path: host pointers in, host pointers out) --------
 * Same semantics as vq_forward + vq_backward_tokens + vq_backward_codebook for token-major fp32
 * data living in (preferably pinned) HOST memory.  Batches of more than 64 Ki tokens are streamed in
 * token chunks: host->device copies, kernels and device->host copies of different chunks overlap on
 * two internal copy streams (full-duplex PCIe), results are identical to the unchunked call (integer
 * partial sums).  Everything is ordered after prior work on `stream`, and `stream` waits for the last
 * copy.  `dev_arena` is device scratch of vq_host_step_arena_bytes().                              */
VQ_API int vq_host_step_arena_bytes(int64_t T, int K, int D, size_t* out);
VQ_API int vq_host_step(const float* z_host, const float* g_zq_host, int64_t T,
                 const float* weight_host, int K, int D, int form, float beta,
                 float* z_q_host, int64_t* idx_host, float* loss_host,
                 float* grad_z_host, float* grad_weight_host, int64_t* stats_host,
                 void* dev_arena, size_t arena_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VQ_B200_H */
