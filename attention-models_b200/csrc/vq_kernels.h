// Internal host-side launchers (one per kernel family).  Not part of the public ABI.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace vq {

// Prepared codebook, carved out of the opaque `cb` blob handed across the ABI.
struct CodebookView {
    float* en32;        // K*D  unit codes, fp32
    float* code_sq;     // K    sum(en^2) in ATen order (~1, 0 for zero codes)
    float* code_denom;  // K    max(||E_k||, eps)
    __half* en16;       // K*D  fp16 copy of en32 (tensor-core operand, TMA source)
    int* info;          // 64 per-block counts of codes whose |en|^2 is not ~1 (zero / non-finite rows); any != 0: degenerate
    // D = 32, K % 512 == 0 only (else null): "cell" copies for the fp16-accumulator search.  A cell is the 8 codes
    // g*512 + hs + 64*m (m = 0..7) that share one slot of group g; cell id ci = g*64 + hs.
    float* en32c;       // K*D  [ci][q][m][4]: 16-byte chunk q of code m -- 8 lanes, one per code, read whole lines
    float* csq_cell;    // K    [ci][m] = code_sq of that code
    int K, D;
    int cell_kind;      // cell_layout_kind(K, D)
};
size_t codebook_bytes(int K, int D);
CodebookView codebook_view(void* cb, int K, int D);

constexpr int kSegPiece = 256;   // tokens per work item of the segmented codebook-gradient sum
constexpr int kCandExactBit = 0x40000000;  // cand[t] holds a final index, not a cell id

// ---- vq_prep.cu ------------------------------------------------------------------------------
cudaError_t launch_prep_codebook(const float* weight, const CodebookView& cb, cudaStream_t s, bool raw = false);
// token-major rows -> zn32, row_sq, denom, zn16 (any output may be null); also clears the buffers of `zl`
struct ZeroList {
    void* ptr[4];
    unsigned long long bytes[4];
};
cudaError_t launch_prep_tokens(const float* z, int64_t T, int D, float* zn32, float* row_sq, float* denom,
                               __half* zn16, const ZeroList& zl, cudaStream_t s, bool raw = false);
// denom[t] = 1 for the un-normalised form behind the NCHW transposition (which then divides by nothing)
cudaError_t launch_fill_ones(float* p, int64_t n, cudaStream_t s);
cudaError_t launch_zero_ranges(const ZeroList& zl, cudaStream_t s);
// codebook + token preparation in one launch (training step; token-major rows, prep_fusable(D))
bool prep_fusable(int D);
cudaError_t launch_prep_fused(const float* weight, const CodebookView& cb, const float* z, int64_t T, float* zn32, float* row_sq,
                              float* denom, __half* zn16, const ZeroList& zl, cudaStream_t s);
// NCHW (b, D, hw): denominators in ATen's channel-strided order (schedule chosen from T, hw, D)
cudaError_t launch_norm_nchw(const float* z, int64_t T, int64_t hw, int D, float* denom, const ZeroList& zl, cudaStream_t s);
// (b, D, hw) -> (T, D), optionally divided by denom[t]; optional fp16 copy
cudaError_t launch_nchw_to_tok(const float* in, int64_t T, int64_t hw, int D, const float* denom,
                               float* out32, __half* out16, cudaStream_t s);
cudaError_t launch_tok_to_nchw(const float* in, int64_t T, int64_t hw, int D, float* out, cudaStream_t s);
// row_sq[t] = sum(zn^2) over contiguous rows, ATen order (used after the NCHW path)
cudaError_t launch_row_sumsq(const float* zn32, int64_t T, int D, float* row_sq, cudaStream_t s);
// the three NCHW prep launches above in one (hw % 4 == 0, D in {64, 128, 256}, ATen's 4-stripe reduction shape)
bool prep_nchw_fused_supported(int64_t T, int64_t hw, int D);
cudaError_t launch_prep_nchw_fused(const float* z, int64_t T, int64_t hw, int D, float* denom, float* zn32, __half* zn16,
                                   float* row_sq, const ZeroList& zl, bool raw, cudaStream_t s);

// ---- vq_prequant.cu (pre_quant / post_quant projection fusion) ------------------------------
// z = x W^T + b (x: (T, C), W: (D, C)) and the token preparation above in one pass; z itself is written only if z_out != null
bool prequant_supported(int C, int D);
cudaError_t launch_prequant_prep(const float* x, int64_t T, int C, const float* w, const float* bias, int D, float* zn32,
                                 float* row_sq, float* denom, __half* zn16, float* z_out, const ZeroList& zl, cudaStream_t s);
// table[k] = W y_k + b over K rows of y (K, D); W: (C, D)
cudaError_t launch_project_codebook(const float* y, int K, int D, const float* w, const float* bias, int C, float* table,
                                    cudaStream_t s);

// ---- vq_dist_simt.cu -------------------------------------------------------------------------
// Exhaustive fp32 search.  rows == nullptr: all T rows; else the first *n_rows entries of `rows`.
// Writes cand[row] = index | kCandExactBit and counts near-tie rows into stats.
// partial_ws (scan_partial_bytes) lets listed rows split the codebook over blocks; may be null.
size_t scan_partial_bytes(int64_t T);
// min_rows: a listed scan leaves at once when the list holds at most min_rows entries (someone else took them)
// tile_done_zeroed: kScanTileCounters ints the caller has already cleared on the stream (else cleared here, by a memset)
constexpr int kScanTileCounters = 256;
cudaError_t launch_scan_exact(const float* zn32, const float* row_sq, const CodebookView& cb, int64_t T,
                              const int* rows, const int* n_rows, int64_t max_rows, int* cand, int64_t* stats,
                              void* partial_ws, cudaStream_t s, int min_rows = 0, int* tile_done_zeroed = nullptr);

// outputs of the finish pass for rows a search kernel finishes itself (all null: search only)
struct ListedFinish {
    float* zq = nullptr; void* idx = nullptr; int32_t* hist = nullptr; unsigned long long* seg = nullptr;
    int idx_bits = 64;
};
cudaError_t launch_scan_listed_tail(const float* zn32, const float* row_sq, const CodebookView& cb, int64_t T,
                                    const int* rows, const int* n_rows, int64_t row_begin, int* cand, int64_t* stats,
                                    const ListedFinish& fin, cudaStream_t s);

// ---- vq_dist_tc16.cu (D = 32: fp16 accumulators, packed 16-bit maxima) ------------------------
bool tc16_supported(int64_t T, int K, int D);
int tc16_max_clusters();      // > 0: the D = 32 filter runs as clusters of two 128-row CTAs (token boxes of 128 rows, code boxes of 64)
constexpr int kFlaggedCap = 4096;     // listed rows that get the sliced per-row search (and a done counter each)
// exact rescoring of the filter's records + sliced search of the listed rows + the finish pass, one launch
cudaError_t launch_exact_finish16(const void* records, const float* zn32, const float* row_sq, const CodebookView& cb, int64_t T,
                                  const int* flagged, const int* n_flagged, int* done_counters, void* partial_ws,
                                  float* zq_tok, void* idx_out, int32_t* hist, int64_t* seg_sums, int64_t* stats,
                                  cudaStream_t s, int idx_bits = 64);

// ---- vq_dist_tc.cu ---------------------------------------------------------------------------
// tcgen05 search: cand[row] = cell id (or exact index for rows resolved in-kernel); rows it cannot
// decide are appended to flagged[] (count in *n_flagged).
bool tc_supported(int64_t T, int K, int D);
size_t tc_workspace_bytes(int64_t T, int K, int D);
int tc_flag_multiplier(int64_t T, int K, int D);   // entries the undecided-row list may need per token row
// generic D: launch_dist_tc is the filter alone; launch_rescore_generic (k_rescore_g) then rescores the decided rows and
// searches the listed rows itself when there are at most kFewFlagged of them (needs n_flagged[64 ..] zeroed: kFewFlagged
// done counters; partial_ws >= kFewFlagged * kFewFlaggedSlices * 16 B); longer lists are for
// launch_scan_exact(..., min_rows = kFewFlagged)
constexpr int kFewFlagged = 128;
constexpr int kFewFlaggedSlices = 256;   // cell slices a listed row's search is cut into there
cudaError_t launch_dist_tc(const __half* zn16, const float* zn32, const float* row_sq, const CodebookView& cb,
                           int64_t T, int* cand, int* flagged, int* n_flagged, int64_t* stats, void* tc_ws,
                           void* partial_ws, cudaStream_t s);
cudaError_t launch_rescore_generic(const float* zn32, const float* row_sq, const CodebookView& cb, int64_t T, int* cand,
                                   const int* flagged, int* n_flagged, int64_t* stats, const void* tc_ws, void* partial_ws,
                                   cudaStream_t s);

// ---- vq_finish.cu ----------------------------------------------------------------------------
// idx / hist / z_q (token-major) / loss partial from final indices in cand[]; with seg_sums (K*D + K int64, not
// zeroed here) also the codebook-gradient segment sums S_k += fixed(q_k - zn_t) as integer reductions.
cudaError_t launch_finish(const float* zn32, const int* cand, const CodebookView& cb, int64_t T,
                          float* zq_tok, void* idx_out, int32_t* hist, int64_t* seg_sums, int64_t* stats, cudaStream_t s,
                          int idx_bits = 64);      // idx_out: int64, or int32 / uint16 (store_token, vq_common.cuh)
// the same with z_q written straight into the (b, D, hw) output
cudaError_t launch_finish_nchw(const float* zn32, const int* cand, const CodebookView& cb, int64_t T, int64_t hw,
                               float* zq_nchw, void* idx_out, int32_t* hist, int64_t* seg_sums, int64_t* stats,
                               cudaStream_t s, int idx_bits = 64);
cudaError_t launch_loss_finalize(const int64_t* stats, int64_t n_elem_total, int form, float beta, float* loss,
                                 cudaStream_t s);
cudaError_t launch_gather(const void* idx, int64_t T, int64_t hw, const float* table, int K, int D,
                          int layout_out, float* out, int64_t* stats, cudaStream_t s, int idx_bits = 64);

// ---- vq_backward.cu --------------------------------------------------------------------------
size_t backward_workspace_bytes(int64_t T, int K, int D);
cudaError_t launch_backward_tokens(const float* g_tok, const float* zn32, const float* denom, const int64_t* idx,
                                   const CodebookView& cb, int64_t T, float coef_commit, const float* g_loss,
                                   float* grad_tok, cudaStream_t s);
// grad_z and grad_E (+ loss) in one launch, for segment sums that already exist (vq_forward's)
cudaError_t launch_backward_fused(const float* g_tok, const float* zn32, const float* denom, const int64_t* idx,
                                  const CodebookView& cb, int64_t T, float coef_commit, const float* g_loss, float* grad_tok,
                                  const int64_t* seg_sums, float coef_codebook, float* grad_weight, const int64_t* stats,
                                  int64_t n_elem_total, int form, float beta, float* loss, cudaStream_t s);
// grad_z for the (b, D, hw) layout: upstream gradient read and grad_z written in place of the two layout kernels
cudaError_t launch_backward_tokens_nchw(const float* g_nchw, const float* zn32, const float* denom, const int64_t* idx,
                                        const CodebookView& cb, int64_t T, int64_t hw, float coef_commit, const float* g_loss,
                                        float* grad_nchw, cudaStream_t s, bool raw = false);
cudaError_t launch_segment_sums(const float* zn32, const int64_t* idx, const int32_t* hist, const CodebookView& cb,
                                int64_t T, int64_t* seg_sums, void* ws, size_t ws_bytes, cudaStream_t s);
// grad_E from the segment sums; with `loss` also the loss from stats (the former k_loss_finalize launch)
cudaError_t launch_codebook_grad(const int64_t* seg_sums, const CodebookView& cb, float coef, const float* g_loss,
                                 float* grad_weight, const int64_t* stats, int64_t n_elem_total, int form, float beta,
                                 float* loss, cudaStream_t s);

// ---- vq_tokens.cu ---------------------------------------------------------------------------
cudaError_t launch_token_embed(const void* tokens, const uint8_t* mask, int64_t T, int64_t n_per_seq, int64_t mask_token_id,
                               int64_t ignore_index, const float* table, int64_t V, int dim, const float* pos, float* embeds,
                               int64_t* input_ids, int64_t* labels, int64_t* stats, cudaStream_t s, const float* start = nullptr,
                               int token_bits = 64);
// token indices between the wire formats (bits in {16, 32, 64}; values that do not fit the target are the caller's error)
cudaError_t launch_tokens_convert(const void* in, int in_bits, void* out, int out_bits, int64_t T, cudaStream_t s);
// grad_table[id] += grad_out[row]: integer-accumulated (deterministic) embedding backward; `normalised` != null applies the
// backward of l2norm(E[k]) on top (grad_table is then the codebook gradient, V == K, dim == D)
size_t embedding_backward_bytes(int64_t V, int dim);
cudaError_t launch_embedding_backward(const int64_t* ids, int64_t n_ids, int64_t ids_per_seq, int64_t rows_per_seq,
                                      int64_t row_shift, const float* grad_out, int64_t n_rows, int64_t hw, int64_t V, int dim,
                                      const CodebookView* normalised, float* grad_table, void* ws, cudaStream_t s);

// ---- vq_peer.cu ------------------------------------------------------------------------------
// byte layout of one rank's exchange buffer (see vq_peer.cu); offsets of stats / hist are relative to the slot
struct ExchangeLayout {
    size_t seg_bytes, stats_off, hist_off, slot_bytes, slot0_off, total;
    size_t results_off, results_hist_off;      // grad_E (K*D fp32) and, results_hist_off behind it, the histogram (K int64)
};
ExchangeLayout exchange_layout(int K, int D);
cudaError_t peer_configure(void* own_buf, unsigned long long timeout_ns, void* abort_flag, cudaStream_t s);
cudaError_t peer_resync(void* own_buf, cudaStream_t s);
// token backward (token-major rows) to run in the same launch, behind the exchange's blocks; null: none
struct TokenBackward {
    const float* g_tok; const float* zn32; const float* denom; const int64_t* idx; int64_t T; float coef_commit; float* grad_tok;
};
cudaError_t launch_codebook_grad_sharded(const void* const* peer_bufs, int world, int rank, int slot, unsigned epoch,
                                         const CodebookView& cb, float coef, const float* g_loss, int64_t n_elem_total,
                                         int form, float beta, float* grad_weight, int64_t* hist_total, float* loss,
                                         int64_t* stats_total, const TokenBackward* tokens, cudaStream_t s);

// dispatch on the supported codebook dims (powers of two in [16, 512])
// Cell copies of the unit codes for the exact rescoring behind a tensor-core filter (CodebookView::en32c):
//   1  D = 32 filter (vq_dist_tc16.cu): cell = the 8 codes g*512 + hs + 64*m, ci = g*64 + hs
//   2  generic filter (vq_dist_tc.cu):  cell = slot s of the 256-code group g, ci = g*32 + s, member i = code
//      g*256 + 64*(i>>1) + 32*(s>>4) + (s&15) + 16*(i&1)
//   0  none (shapes the tensor-core search does not take)
bool tc16_wide_drain();     // vq_dist_tc16.cu: VQ_TC16_W16=1 selects the 16-epilogue-warp kernel (cell layout 3)
bool tc16_disabled();       // VQ_TC16_DISABLE=1: D = 32 goes through the generic filter (cell layout 2) like every other D
inline int cell_layout_kind(int K, int D) {
    if (D == 32 && K >= 512 && (K % 512) == 0 && K <= 65536 && !tc16_disabled()) return tc16_wide_drain() ? 3 : 1;
    if ((D == 32 || D == 64 || D == 128 || D == 256) && K >= 256 && (K % 256) == 0) return 2;
    return 0;
}
// D = 32 filter cells (kinds 1 and 3): cell id ci = group * 64 + cell-in-group, member m = 0..7.
//   kind 1 (8 epilogue warps, a thread folds column c with c + 64):  code = g*512 + (ci & 63) + 64*m
//   kind 3 (16 epilogue warps, a thread owns a 64-column half h2 of every tile and folds column c with c + 32):
//           ci & 63 = h2*32 + s,  code = g*512 + 128*(m >> 1) + 64*h2 + 32*(m & 1) + s
__host__ __device__ inline int tc16_code_of(int ci, int m, int kind) {
    const int g = ci >> 6, c = ci & 63;
    if (kind == 1) return g * 512 + c + 64 * m;
    return g * 512 + 128 * (m >> 1) + 64 * (c >> 5) + 32 * (m & 1) + (c & 31);
}
__host__ __device__ inline void tc16_cell_of(int code, int kind, int& ci, int& m) {
    const int g = code >> 9, w = code & 511;
    if (kind == 1) { ci = g * 64 + (w & 63); m = w >> 6; return; }
    const int t = w >> 7, c = w & 127;
    ci = g * 64 + (c >> 6) * 32 + (c & 31);
    m = 2 * t + ((c >> 5) & 1);
}
inline bool has_cell_layout(int K, int D) { return cell_layout_kind(K, D) != 0; }
// (cell, member) of a code in the generic layout
__host__ __device__ inline void generic_cell_of(int code, int& ci, int& member) {
    const int g = code >> 8, w = code & 255;
    ci = g * 32 + (((w >> 5) & 1) << 4) + (w & 15);
    member = ((w >> 6) << 1) | ((w >> 4) & 1);
}
inline bool dim_supported(int D) { return D >= 16 && D <= 512 && (D & (D - 1)) == 0; }

#define VQ_DISPATCH_D(D, ...)                                       \
    switch (D) {                                                    \
        case 16:  { constexpr int kD = 16;  __VA_ARGS__; } break;   \
        case 32:  { constexpr int kD = 32;  __VA_ARGS__; } break;   \
        case 64:  { constexpr int kD = 64;  __VA_ARGS__; } break;   \
        case 128: { constexpr int kD = 128; __VA_ARGS__; } break;   \
        case 256: { constexpr int kD = 256; __VA_ARGS__; } break;   \
        case 512: { constexpr int kD = 512; __VA_ARGS__; } break;   \
        default: return cudaErrorInvalidValue;                      \
    }

int sm_count();          // of the CURRENT device (cached per device)

// Function attributes (opt-in shared memory sizes) are per device: a process that uses several GPUs has to set them on
// each.  `need()` returns true the first time it is called on the current device.
struct PerDeviceOnce {
    unsigned long long done = 0;                 // bit d: configured on device d (d < 64; beyond that: set every time)
    bool need() {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
        const unsigned long long bit = 1ull << dev;
        const unsigned long long before = __atomic_fetch_or(&done, bit, __ATOMIC_RELAXED);
        return (before & bit) == 0;
    }
};

// Programmatic dependent launch (PDL) for the kernels of the step: a kernel launched with launch_pdl may be
// scheduled while its predecessor in the stream drains (that predecessor calls pdl_trigger()), and must call
// pdl_wait() before it touches anything an earlier kernel wrote -- the wait returns when the predecessor grid has
// completed and flushed.  Every such kernel waits, so completion stays transitive along the stream; kernels launched
// the classic way are unaffected.  Off by default: with 5 launches per step the measured gain is < 1 us of 245 us
// (the stream has no gaps to hide once the launches are queued ahead), and overlapping grids blur per-kernel
// timings; VQ_PDL=1 in the environment turns the attribute on.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// ... as thread-block clusters of `cluster` CTAs along x (grid.x a multiple of it)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, int cluster,
                                      Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// kernel-launch counter (bench.py's gpu_launches); bumped by every launcher next to its <<<>>>
extern long long g_kernel_launches;
inline void count_launch(int n = 1) { __atomic_fetch_add(&g_kernel_launches, (long long)n, __ATOMIC_RELAXED); }

}  // namespace vq
