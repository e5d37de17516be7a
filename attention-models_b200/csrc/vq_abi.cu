// extern "C" entry points of libvq_b200.so (declared in include/vq_b200.h).
// Orchestration only: argument checks, workspace carving, kernel sequencing on the caller's stream.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "../../include/vq_b200.h"
#include "vq_kernels.h"

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

int cuda_fail(cudaError_t e, const char* what) {
    return fail(VQ_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

#define VQ_CUDA(call)                                             \
    do {                                                          \
        cudaError_t e__ = (call);                                 \
        if (e__ != cudaSuccess) return cuda_fail(e__, #call);     \
    } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// measurement hooks: CUDA-event pairs around the kernels of a step while profiling is on.  Event records between
// kernels cost a few microseconds per step, so only every g_profile_every-th vq_forward (and the backward calls that
// follow it) is bracketed; the other steps of the timed region run undisturbed.
bool g_profiling = false, g_sampled = false;
int g_profile_every = 1;
unsigned g_profile_mask = ~0u;
long long g_profile_calls = 0;
long long g_launches_at_begin = 0;
std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_slot_events[VQ_PROFILE_SLOTS];

struct SlotTimer {
    cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t s;
    int slot;
    SlotTimer(cudaStream_t stream, int which) : s(stream), slot(which) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (g_profiling && g_sampled && cudaStreamIsCapturing(stream, &cap) == cudaSuccess &&
            cap != cudaStreamCaptureStatusNone)
            return;                                    // events recorded into a graph cannot be read back: no hooks
        if (g_profiling && g_sampled && ((g_profile_mask >> which) & 1u) && cudaEventCreate(&a) == cudaSuccess && cudaEventCreate(&b) == cudaSuccess) cudaEventRecord(a, s);
    }
    void stop() {
        if (a && b) { cudaEventRecord(b, s); g_slot_events[slot].emplace_back(a, b); a = b = nullptr; }
    }
};

struct Bump {
    char* base; size_t off = 0;
    explicit Bump(void* p) : base(static_cast<char*>(p)) {}
    template <typename T> T* take(size_t n) {
        T* r = reinterpret_cast<T*>(base ? base + off : nullptr);
        off += align_up(sizeof(T) * n, 256);
        return r;
    }
};

struct FwdWs {
    float* zn32; float* denom; float* row_sq; __half* zn16; int* cand; int* flagged; int* n_flagged;
    int64_t* stats; float* zq_tok; void* tc_ws; void* scan_ws; size_t bytes;
};

// Sized for the worst case over the optional outputs so one size serves every flag combination.
FwdWs carve_forward(void* ws, int64_t T, int K, int D) {
    Bump b(ws);
    FwdWs w;
    const size_t n = (size_t)(T > 0 ? T : 1);
    w.zn32 = b.take<float>(n * D);
    w.denom = b.take<float>(n);
    w.row_sq = b.take<float>(n);
    w.zn16 = b.take<__half>(n * D);
    w.cand = b.take<int>(n);
    w.flagged = b.take<int>(n * (size_t)vq::tc_flag_multiplier(T, K, D));
    w.n_flagged = b.take<int>(64 + vq::kFlaggedCap);   // [0]: count, [64..]: done counters of the sliced fallback
    w.stats = b.take<int64_t>(VQ_STATS_LEN);
    w.zq_tok = b.take<float>(n * D);
    const size_t tcb = vq::tc_workspace_bytes(T, K, D);
    w.tc_ws = b.take<char>(tcb ? tcb : 1);
    w.scan_ws = b.take<char>(vq::scan_partial_bytes(T));
    w.bytes = b.off;
    return w;
}

int check_dims(int64_t T, int K, int D) {
    if (!vq::dim_supported(D)) return fail(VQ_ERR_ARG, "codebook_dim %d unsupported (powers of two in [16, 512])", D);
    if (K <= 0 || K >= (1 << 30)) return fail(VQ_ERR_ARG, "codebook_size %d out of range", K);
    if (T < 0 || T >= (1ll << 31)) return fail(VQ_ERR_ARG, "token count %lld out of range", (long long)T);
    return VQ_OK;
}

int check_layout(int layout, int64_t T, int64_t hw) {
    if (layout == VQ_LAYOUT_TOKEN_MAJOR) return VQ_OK;
    if (layout != VQ_LAYOUT_NCHW) return fail(VQ_ERR_ARG, "unknown layout %d", layout);
    if (hw <= 0 || T % hw != 0) return fail(VQ_ERR_ARG, "NCHW layout needs hw > 0 dividing T (T=%lld hw=%lld)",
                                            (long long)T, (long long)hw);
    return VQ_OK;
}

}  // namespace

extern "C" {

int vq_abi_version(void) { return VQ_ABI_VERSION; }

const char* vq_last_error(void) { return g_last_error.c_str(); }

int vq_uses_tensor_cores(int64_t T, int K, int D) {
    if (check_dims(T, K, D) != VQ_OK) return 0;
    return vq::tc_supported(T, K, D) ? 1 : 0;
}

int vq_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    cudaDeviceProp p;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
        cudaGetLastError();
        if (sm_count) *sm_count = 0;
        if (cc_major) *cc_major = 0;
        if (cc_minor) *cc_minor = 0;
        return fail(VQ_ERR_CUDA, "no CUDA device");
    }
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return VQ_OK;
}

int vq_codebook_bytes(int K, int D, size_t* out) {
    if (!out) return fail(VQ_ERR_ARG, "out is NULL");
    if (int r = check_dims(0, K, D)) return r;
    *out = vq::codebook_bytes(K, D);
    return VQ_OK;
}

int vq_codebook_prepare(const float* weight, int K, int D, void* cb, size_t cb_bytes, void* stream) {
    if (int r = check_dims(0, K, D)) return r;
    if (!weight || !cb) return fail(VQ_ERR_ARG, "weight/cb is NULL");
    if (cb_bytes < vq::codebook_bytes(K, D)) return fail(VQ_ERR_WORKSPACE, "codebook blob too small");
    vq::CodebookView v = vq::codebook_view(cb, K, D);
    if (g_profiling) g_sampled = (g_profile_calls % g_profile_every) == 0;     // the step the next vq_forward opens
    SlotTimer timer(static_cast<cudaStream_t>(stream), VQ_PROFILE_PREP_CODEBOOK);
    VQ_CUDA(vq::launch_prep_codebook(weight, v, static_cast<cudaStream_t>(stream)));
    timer.stop();
    return VQ_OK;
}

int vq_codebook_prepare_raw(const float* weight, int K, int D, void* cb, size_t cb_bytes, void* stream) {
    if (int r = check_dims(0, K, D)) return r;
    if (!weight || !cb) return fail(VQ_ERR_ARG, "weight/cb is NULL");
    if (cb_bytes < vq::codebook_bytes(K, D)) return fail(VQ_ERR_WORKSPACE, "codebook blob too small");
    VQ_CUDA(vq::launch_prep_codebook(weight, vq::codebook_view(cb, K, D), static_cast<cudaStream_t>(stream), true));
    return VQ_OK;
}

int vq_workspace_bytes(int64_t T, int K, int D, int flags, size_t* out) {
    (void)flags;
    if (!out) return fail(VQ_ERR_ARG, "out is NULL");
    if (int r = check_dims(T, K, D)) return r;
    *out = carve_forward(nullptr, T, K, D).bytes;
    return VQ_OK;
}

namespace {
// pre_quant projection in front of the quantiser (vq_forward_projected): z = x W^T + b is formed inside the token prep
struct Projection {
    const float* x; int C; const float* w; const float* bias; float* z_out;
};

int forward_impl(const float* z, const Projection* pj, int layout, int64_t T, int64_t hw, const float* weight, void* cb, int K,
                 int D, int form, float beta, int flags, int64_t n_elem_total, float* z_q, void* idx, float* loss, int32_t* hist,
                 int64_t* stats, float* saved_zn, float* saved_denom, int64_t* seg_sums, void* ws, size_t ws_bytes, void* stream);
}  // namespace

int vq_forward(const float* z, int layout, int64_t T, int64_t hw, const float* weight, void* cb, int K, int D, int form, float beta,
               int flags, int64_t n_elem_total, float* z_q, void* idx, float* loss, int32_t* hist, int64_t* stats,
               float* saved_zn, float* saved_denom, int64_t* seg_sums, void* ws, size_t ws_bytes, void* stream) {
    return forward_impl(z, nullptr, layout, T, hw, weight, cb, K, D, form, beta, flags, n_elem_total, z_q, idx, loss, hist, stats,
                        saved_zn, saved_denom, seg_sums, ws, ws_bytes, stream);
}

int vq_prequant_supported(int C, int D) { return vq::prequant_supported(C, D) ? 1 : 0; }

int vq_forward_projected(const float* x, int C, const float* w_pre, const float* b_pre, int64_t T, const float* weight, void* cb,
                         int K, int D, int form, float beta, int flags, int64_t n_elem_total, float* z_q, void* idx, float* loss,
                         int32_t* hist, int64_t* stats, float* saved_zn, float* saved_denom, int64_t* seg_sums, float* z_out,
                         void* ws, size_t ws_bytes, void* stream) {
    if (form != VQ_FORM_VIT) return fail(VQ_ERR_ARG, "vq_forward_projected: the Linear pre_quant belongs to the ViT form (form %d)", form);
    if (!vq::prequant_supported(C, D))
        return fail(VQ_ERR_ARG, "vq_forward_projected: in_features %d -> codebook_dim %d unsupported (D = 32, C a multiple of 64 "
                                "up to 768); project with a GEMM and call vq_forward", C, D);
    if ((!x && T > 0) || !w_pre) return fail(VQ_ERR_ARG, "x/w_pre is NULL");
    if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) return fail(VQ_ERR_ARG, "x must be 16-byte aligned");
    const Projection pj{x, C, w_pre, b_pre, z_out};
    return forward_impl(nullptr, &pj, VQ_LAYOUT_TOKEN_MAJOR, T, 0, weight, cb, K, D, form, beta, flags, n_elem_total, z_q, idx, loss,
                        hist, stats, saved_zn, saved_denom, seg_sums, ws, ws_bytes, stream);
}

int vq_project_codebook(const float* weight, const void* cb, int K, int D, int normalise, const float* w_post, const float* b_post,
                        int C, float* table, void* stream) {
    if (int r = check_dims(0, K, D)) return r;
    if (C <= 0 || !w_post || !table) return fail(VQ_ERR_ARG, "w_post/table is NULL or out_features %d <= 0", C);
    if ((reinterpret_cast<uintptr_t>(w_post) & 15) != 0) return fail(VQ_ERR_ARG, "w_post must be 16-byte aligned");
    const float* y = nullptr;
    if (normalise) {
        if (!cb) return fail(VQ_ERR_ARG, "normalise=1 needs the prepared codebook");
        y = vq::codebook_view(const_cast<void*>(cb), K, D).en32;
    } else {
        if (!weight) return fail(VQ_ERR_ARG, "normalise=0 needs the raw weight");
        y = weight;
    }
    VQ_CUDA(vq::launch_project_codebook(y, K, D, w_post, b_post, C, table, static_cast<cudaStream_t>(stream)));
    return VQ_OK;
}

int vq_gather_projected(const void* tokens, int token_bits, int64_t T, int64_t hw, const float* table, int K, int C, int layout_out,
                        float* out, int64_t* stats, void* stream) {
    if (K <= 0 || K >= (1 << 30)) return fail(VQ_ERR_ARG, "codebook_size %d out of range", K);
    if (T < 0 || T >= (1ll << 31)) return fail(VQ_ERR_ARG, "token count %lld out of range", (long long)T);
    if (C <= 0 || (layout_out == VQ_LAYOUT_TOKEN_MAJOR && C % 4 != 0))
        return fail(VQ_ERR_ARG, "out_features %d: positive, and a multiple of 4 for token-major output", C);
    if (int r = check_layout(layout_out, T, hw)) return r;
    if (token_bits != 16 && token_bits != 32 && token_bits != 64) return fail(VQ_ERR_ARG, "token_bits %d: 16, 32 or 64", token_bits);
    if (!tokens || !out || !table) return fail(VQ_ERR_ARG, "tokens/table/out is NULL");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (stats) VQ_CUDA(cudaMemsetAsync(stats + VQ_STAT_BAD_INDEX, 0, sizeof(int64_t), s));
    VQ_CUDA(vq::launch_gather(tokens, T, hw, table, K, C, layout_out, out, stats, s, token_bits));
    return VQ_OK;
}

namespace {
int forward_impl(const float* z, const Projection* pj, int layout, int64_t T, int64_t hw, const float* weight, void* cb, int K,
                 int D, int form, float beta, int flags, int64_t n_elem_total, float* z_q, void* idx, float* loss, int32_t* hist,
                 int64_t* stats, float* saved_zn, float* saved_denom, int64_t* seg_sums, void* ws, size_t ws_bytes, void* stream) {
    if (int r = check_dims(T, K, D)) return r;
    if (int r = check_layout(layout, T, hw)) return r;
    if (form != VQ_FORM_VIT && form != VQ_FORM_VQGAN && form != VQ_FORM_VQGAN_L2) return fail(VQ_ERR_ARG, "unknown form %d", form);
    const bool raw = form == VQ_FORM_VQGAN_L2;
    if (raw) flags |= VQ_FLAG_EXACT_SCAN;       // the filters' error bounds assume unit rows
    if (!cb || !idx || (!pj && !z && T > 0)) return fail(VQ_ERR_ARG, "z/cb/idx is NULL");
    const bool indices_only = (flags & VQ_FLAG_INDICES_ONLY) != 0;
    if (!indices_only && !z_q) return fail(VQ_ERR_ARG, "z_q is NULL without VQ_FLAG_INDICES_ONLY");
    if (indices_only && seg_sums) return fail(VQ_ERR_ARG, "seg_sums needs the full forward (no VQ_FLAG_INDICES_ONLY)");
    // narrow token formats: encode-only calls (the backward reads the int64 indices of a full forward)
    if ((flags & VQ_FLAG_IDX32) && (flags & VQ_FLAG_IDX16)) return fail(VQ_ERR_ARG, "VQ_FLAG_IDX32 and VQ_FLAG_IDX16 exclude each other");
    const int idx_bits = (flags & VQ_FLAG_IDX16) ? 16 : ((flags & VQ_FLAG_IDX32) ? 32 : 64);
    if (idx_bits != 64 && !indices_only) return fail(VQ_ERR_ARG, "narrow indices need VQ_FLAG_INDICES_ONLY");
    if (idx_bits == 16 && K > 65536) return fail(VQ_ERR_ARG, "VQ_FLAG_IDX16 needs codebook_size <= 65536 (K=%d)", K);
    if (!ws) return fail(VQ_ERR_WORKSPACE, "workspace is NULL");
    FwdWs w = carve_forward(ws, T, K, D);
    if (ws_bytes < w.bytes) return fail(VQ_ERR_WORKSPACE, "workspace too small: %zu < %zu", ws_bytes, w.bytes);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    vq::CodebookView cbv = vq::codebook_view(cb, K, D);
    if (g_profiling) g_sampled = (g_profile_calls++ % g_profile_every) == 0;
    // `weight` given: the codebook is prepared by this call -- in the token prep launch when the shape allows it
    const bool fuse_prep = weight && layout == VQ_LAYOUT_TOKEN_MAJOR && vq::prep_fusable(D) && !raw && !pj;
    if (weight && !fuse_prep) {
        SlotTimer cb_timer(s, VQ_PROFILE_PREP_CODEBOOK);
        VQ_CUDA(vq::launch_prep_codebook(weight, cbv, s, raw));
        cb_timer.stop();
    }

    float* zn32 = saved_zn ? saved_zn : w.zn32;
    float* denom = saved_denom ? saved_denom : w.denom;
    int64_t* st = stats ? stats : w.stats;
    const bool use_tc = !(flags & VQ_FLAG_EXACT_SCAN) && vq::tc_supported(T, K, D);
    __half* zn16 = use_tc ? w.zn16 : nullptr;
    const bool tc16 = use_tc && vq::tc16_supported(T, K, D);

    // 0. buffers the call accumulates into, cleared by the first kernel (no memset nodes)
    vq::ZeroList zl = {};
    int nz = 0;
    auto zero = [&](void* p, size_t bytes) { zl.ptr[nz] = p; zl.bytes[nz] = bytes; ++nz; };
    // [0]: listed rows, [64 ..]: done counters of the listed-row searches, then (generic filter) the tiled scan's tile counters
    if (use_tc) zero(w.n_flagged, sizeof(int) * (tc16 ? 64 + vq::kFlaggedCap : 64 + vq::kFewFlagged + vq::kScanTileCounters));
    if (!(flags & VQ_FLAG_KEEP_STATS)) {
        zero(st, sizeof(int64_t) * VQ_STATS_LEN);
        if (hist) zero(hist, sizeof(int32_t) * (size_t)K);
        if (seg_sums) zero(seg_sums, sizeof(int64_t) * ((size_t)K * D + K));
    }

    // 1. unit rows (ATen-order norms), fp16 copy for the tensor cores
    SlotTimer prep_timer(s, VQ_PROFILE_PREP_TOKENS);
    if (pj) {
        // pre_quant Linear + token prep in one pass over the encoder rows (vq_prequant.cu)
        VQ_CUDA(vq::launch_prequant_prep(pj->x, T, pj->C, pj->w, pj->bias, D, zn32, w.row_sq, denom, zn16, pj->z_out, zl, s));
    } else if (fuse_prep) {
        VQ_CUDA(vq::launch_prep_fused(weight, cbv, z, T, zn32, w.row_sq, denom, zn16, zl, s));
    } else if (layout == VQ_LAYOUT_TOKEN_MAJOR) {
        VQ_CUDA(vq::launch_prep_tokens(z, T, D, zn32, w.row_sq, denom, zn16, zl, s, raw));
    } else if (vq::prep_nchw_fused_supported(T, hw, D)) {
        VQ_CUDA(vq::launch_prep_nchw_fused(z, T, hw, D, denom, zn32, zn16, w.row_sq, zl, raw, s));
    } else if (raw) {
        // un-normalised form: the rows as they are (transposed to token-major), denominators of 1, row_sq = sum(z^2)
        if (nz) VQ_CUDA(vq::launch_zero_ranges(zl, s));
        VQ_CUDA(vq::launch_fill_ones(denom, T, s));
        VQ_CUDA(vq::launch_nchw_to_tok(z, T, hw, D, nullptr, zn32, zn16, s));
        VQ_CUDA(vq::launch_row_sumsq(zn32, T, D, w.row_sq, s));
    } else {
        if (T == 0 && nz) VQ_CUDA(vq::launch_zero_ranges(zl, s));
        VQ_CUDA(vq::launch_norm_nchw(z, T, hw, D, denom, zl, s));
        VQ_CUDA(vq::launch_nchw_to_tok(z, T, hw, D, denom, zn32, zn16, s));
        VQ_CUDA(vq::launch_row_sumsq(zn32, T, D, w.row_sq, s));
    }
    prep_timer.stop();

    // 2. nearest code per row, 3. idx, hist, z_q, loss partial
    float* zq_tok = indices_only ? nullptr : (layout == VQ_LAYOUT_NCHW ? w.zq_tok : z_q);
    SlotTimer timer(s, VQ_PROFILE_SEARCH);
    if (tc16) {
        // D = 32: tensor-core filter -> records; one kernel then does the exact rescoring, the search of the undecided
        // rows (sliced over blocks for the first kFlaggedCap of them, one block per row beyond) and the finish pass.
        VQ_CUDA(vq::launch_dist_tc(zn16, zn32, w.row_sq, cbv, T, w.cand, w.flagged, w.n_flagged, st, w.tc_ws, w.scan_ws, s));
        timer.stop();
        SlotTimer exact_timer(s, VQ_PROFILE_EXACT_FINISH);
        VQ_CUDA(vq::launch_exact_finish16(w.tc_ws, zn32, w.row_sq, cbv, T, w.flagged, w.n_flagged, w.n_flagged + 64, w.scan_ws,
                                          zq_tok, idx, hist, seg_sums, st, s, idx_bits));
        exact_timer.stop();
    } else {
        // (profile slots: SEARCH = the tensor-core filter alone, as at D = 32; EXACT_FINISH = everything behind it.
        //  Without the filter the exhaustive scan is the search.)
        if (use_tc) {
            // filter -> exact rescoring (+ the undecided rows when they are few) -> tiled exhaustive scan of a long list
            VQ_CUDA(vq::launch_dist_tc(zn16, zn32, w.row_sq, cbv, T, w.cand, w.flagged, w.n_flagged, st, w.tc_ws, w.scan_ws, s));
            timer.stop();
        } else {
            VQ_CUDA(vq::launch_scan_exact(zn32, w.row_sq, cbv, T, nullptr, nullptr, T, w.cand, st, nullptr, s));
            timer.stop();
        }
        SlotTimer finish_timer(s, VQ_PROFILE_EXACT_FINISH);
        if (use_tc) {
            VQ_CUDA(vq::launch_rescore_generic(zn32, w.row_sq, cbv, T, w.cand, w.flagged, w.n_flagged, st, w.tc_ws, w.scan_ws, s));
            VQ_CUDA(vq::launch_scan_exact(zn32, w.row_sq, cbv, T, w.flagged, w.n_flagged, T * vq::tc_flag_multiplier(T, K, D),
                                          w.cand, st, w.scan_ws, s, vq::kFewFlagged, w.n_flagged + 64 + vq::kFewFlagged));
        }
        if (!indices_only && layout == VQ_LAYOUT_NCHW) {
            VQ_CUDA(vq::launch_finish_nchw(zn32, w.cand, cbv, T, hw, z_q, idx, hist, seg_sums, st, s, idx_bits));
            zq_tok = nullptr;      // written in place: no layout kernel behind
        } else {
            VQ_CUDA(vq::launch_finish(zn32, w.cand, cbv, T, zq_tok, idx, hist, seg_sums, st, s, idx_bits));
        }
        finish_timer.stop();
    }
    if (!indices_only && layout == VQ_LAYOUT_NCHW && zq_tok) VQ_CUDA(vq::launch_tok_to_nchw(zq_tok, T, hw, D, z_q, s));
    if (loss && !indices_only) {
        if (n_elem_total <= 0) return fail(VQ_ERR_ARG, "n_elem_total must be positive");
        VQ_CUDA(vq::launch_loss_finalize(st, n_elem_total, form, beta, loss, s));
    }
    return VQ_OK;
}
}  // namespace

int vq_loss_finalize(const int64_t* stats, int64_t n_elem_total, int form, float beta, float* loss, void* stream) {
    if (!stats || !loss || n_elem_total <= 0) return fail(VQ_ERR_ARG, "bad argument to vq_loss_finalize");
    VQ_CUDA(vq::launch_loss_finalize(stats, n_elem_total, form, beta, loss, static_cast<cudaStream_t>(stream)));
    return VQ_OK;
}

int vq_backward_workspace_bytes(int64_t T, int K, int D, size_t* out) {
    if (!out) return fail(VQ_ERR_ARG, "out is NULL");
    if (int r = check_dims(T, K, D)) return r;
    // segment bucketing only (the NCHW layout needs no token-major staging any more)
    *out = align_up(vq::backward_workspace_bytes(T, K, D), 256);
    return VQ_OK;
}

int vq_backward_tokens(const float* g_zq, int layout, int64_t T, int64_t hw, const float* saved_zn,
                       const float* saved_denom, const int64_t* idx, const int32_t* hist, const void* cb, int K, int D, int form,
                       float beta, const float* g_loss, int64_t n_elem_total, float* grad_z, int64_t* seg_sums,
                       void* ws, size_t ws_bytes, void* stream) {
    if (int r = check_dims(T, K, D)) return r;
    if (int r = check_layout(layout, T, hw)) return r;
    if (!saved_zn || !saved_denom || !idx || !cb) return fail(VQ_ERR_ARG, "saved_zn/saved_denom/idx/cb is NULL");
    if (n_elem_total <= 0) return fail(VQ_ERR_ARG, "n_elem_total must be positive");
    size_t need = 0;
    vq_backward_workspace_bytes(T, K, D, &need);
    if (!ws || ws_bytes < need) return fail(VQ_ERR_WORKSPACE, "backward workspace too small: %zu < %zu", ws_bytes, need);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    vq::CodebookView cbv = vq::codebook_view(const_cast<void*>(cb), K, D);
    char* p = static_cast<char*>(ws);
    void* seg_ws = p;
    const size_t seg_bytes = align_up(vq::backward_workspace_bytes(T, K, D), 256);

    SlotTimer bwd_timer(s, VQ_PROFILE_BACKWARD_TOKENS);
    if (grad_z) {
        const float c1 = (form == VQ_FORM_VIT) ? beta : 1.f;
        const float coef = (float)((double)c1 * 2.0 / (double)n_elem_total);
        if (layout == VQ_LAYOUT_TOKEN_MAJOR) {
            if (form == VQ_FORM_VQGAN_L2) return fail(VQ_ERR_ARG, "VQ_FORM_VQGAN_L2 backward supports the NCHW layout only");
            VQ_CUDA(vq::launch_backward_tokens(g_zq, saved_zn, saved_denom, idx, cbv, T, coef, g_loss, grad_z, s));
        } else {
            VQ_CUDA(vq::launch_backward_tokens_nchw(g_zq, saved_zn, saved_denom, idx, cbv, T, hw, coef, g_loss, grad_z, s,
                                                    form == VQ_FORM_VQGAN_L2));
        }
    }
    if (seg_sums) VQ_CUDA(vq::launch_segment_sums(saved_zn, idx, hist, cbv, T, seg_sums, seg_ws, seg_bytes, s));
    bwd_timer.stop();
    return VQ_OK;
}

int vq_backward_codebook(const int64_t* seg_sums, const void* cb, int K, int D, int form, float beta,
                         const float* g_loss, int64_t n_elem_total, float* grad_weight, const int64_t* stats, float* loss,
                         void* stream) {
    if (int r = check_dims(0, K, D)) return r;
    if (!seg_sums || !cb || !grad_weight || n_elem_total <= 0) return fail(VQ_ERR_ARG, "bad argument to vq_backward_codebook");
    vq::CodebookView cbv = vq::codebook_view(const_cast<void*>(cb), K, D);
    const float c2 = (form == VQ_FORM_VIT) ? 1.f : beta;
    const float coef = (float)((double)c2 * 2.0 / (double)n_elem_total);
    if (loss && !stats) return fail(VQ_ERR_ARG, "loss needs stats");
    SlotTimer timer(static_cast<cudaStream_t>(stream), VQ_PROFILE_CODEBOOK_GRAD);
    VQ_CUDA(vq::launch_codebook_grad(seg_sums, cbv, coef, g_loss, grad_weight, stats, n_elem_total, form, beta, loss,
                                     static_cast<cudaStream_t>(stream)));
    timer.stop();
    return VQ_OK;
}

int vq_backward(const float* g_zq, int layout, int64_t T, int64_t hw, const float* saved_zn, const float* saved_denom,
                const int64_t* idx, const void* cb, int K, int D, int form, float beta, const float* g_loss,
                int64_t n_elem_total, const int64_t* seg_sums, const int64_t* stats, float* grad_z, float* grad_weight,
                float* loss, void* ws, size_t ws_bytes, void* stream) {
    if (!seg_sums || !grad_weight) return fail(VQ_ERR_ARG, "vq_backward needs seg_sums (from vq_forward) and grad_weight");
    if (form == VQ_FORM_VQGAN_L2 && layout == VQ_LAYOUT_TOKEN_MAJOR && grad_z)
        return fail(VQ_ERR_ARG, "VQ_FORM_VQGAN_L2 backward supports the NCHW layout only");
    if (loss && !stats) return fail(VQ_ERR_ARG, "loss needs stats");
    if (grad_z && layout == VQ_LAYOUT_TOKEN_MAJOR && T > 0) {
        if (int r = check_dims(T, K, D)) return r;
        if (!saved_zn || !saved_denom || !idx || !cb) return fail(VQ_ERR_ARG, "saved_zn/saved_denom/idx/cb is NULL");
        if (n_elem_total <= 0) return fail(VQ_ERR_ARG, "n_elem_total must be positive");
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        vq::CodebookView cbv = vq::codebook_view(const_cast<void*>(cb), K, D);
        const float c1 = (form == VQ_FORM_VIT) ? beta : 1.f, c2 = (form == VQ_FORM_VIT) ? 1.f : beta;
        const float coef1 = (float)((double)c1 * 2.0 / (double)n_elem_total), coef2 = (float)((double)c2 * 2.0 / (double)n_elem_total);
        SlotTimer timer(s, VQ_PROFILE_BACKWARD_TOKENS);
        VQ_CUDA(vq::launch_backward_fused(g_zq, saved_zn, saved_denom, idx, cbv, T, coef1, g_loss, grad_z, seg_sums, coef2,
                                          grad_weight, stats, n_elem_total, form, beta, loss, s));
        timer.stop();
        return VQ_OK;
    }
    // other layouts (or no grad_z wanted): the two calls back to back
    if (grad_z)
        if (int r = vq_backward_tokens(g_zq, layout, T, hw, saved_zn, saved_denom, idx, nullptr, cb, K, D, form, beta, g_loss,
                                       n_elem_total, grad_z, nullptr, ws, ws_bytes, stream))
            return r;
    return vq_backward_codebook(seg_sums, cb, K, D, form, beta, g_loss, n_elem_total, grad_weight, stats, loss, stream);
}

int vq_exchange_bytes(int K, int D, size_t* out) {
    if (!out) return fail(VQ_ERR_ARG, "out is NULL");
    if (int r = check_dims(0, K, D)) return r;
    *out = vq::exchange_layout(K, D).total;
    return VQ_OK;
}

int vq_exchange_slot(void* exchange_buf, int K, int D, int slot, int64_t** seg_sums, int64_t** stats, int32_t** hist) {
    if (int r = check_dims(0, K, D)) return r;
    if (!exchange_buf || (slot != 0 && slot != 1)) return fail(VQ_ERR_ARG, "bad exchange buffer / slot");
    const vq::ExchangeLayout L = vq::exchange_layout(K, D);
    char* base = static_cast<char*>(exchange_buf) + L.slot0_off + (size_t)slot * L.slot_bytes;
    if (seg_sums) *seg_sums = reinterpret_cast<int64_t*>(base);
    if (stats) *stats = reinterpret_cast<int64_t*>(base + L.stats_off);
    if (hist) *hist = reinterpret_cast<int32_t*>(base + L.hist_off);
    return VQ_OK;
}

int vq_peer_alloc(size_t bytes, void** dev_ptr, void* ipc_handle_out) {
    if (!dev_ptr || !ipc_handle_out || bytes == 0) return fail(VQ_ERR_ARG, "bad argument to vq_peer_alloc");
    static_assert(sizeof(cudaIpcMemHandle_t) == VQ_IPC_HANDLE_BYTES, "IPC handle size");
    void* p = nullptr;
    VQ_CUDA(cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "vq_peer_alloc"); }
    memcpy(ipc_handle_out, &h, sizeof(h));
    *dev_ptr = p;
    return VQ_OK;
}

int vq_peer_open(const void* ipc_handle, void** dev_ptr) {
    if (!ipc_handle || !dev_ptr) return fail(VQ_ERR_ARG, "bad argument to vq_peer_open");
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, sizeof(h));
    VQ_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return VQ_OK;
}

int vq_peer_close(void* dev_ptr) {
    if (dev_ptr) VQ_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return VQ_OK;
}

int vq_peer_configure(void* own_buf, int64_t timeout_ms, void* abort_flag, void* stream) {
    if (!own_buf || timeout_ms < 0) return fail(VQ_ERR_ARG, "bad argument to vq_peer_configure");
    VQ_CUDA(vq::peer_configure(own_buf, (unsigned long long)timeout_ms * 1000000ull, abort_flag, static_cast<cudaStream_t>(stream)));
    return VQ_OK;
}

int vq_peer_resync(void* own_buf, void* stream) {
    if (!own_buf) return fail(VQ_ERR_ARG, "bad argument to vq_peer_resync");
    VQ_CUDA(vq::peer_resync(own_buf, static_cast<cudaStream_t>(stream)));
    return VQ_OK;
}

int vq_peer_free(void* dev_ptr) {
    if (dev_ptr) VQ_CUDA(cudaFree(dev_ptr));
    return VQ_OK;
}

int vq_backward_codebook_sharded(const void* const* peer_bufs, int world, int rank, int slot, uint32_t epoch, const void* cb,
                                 int K, int D, int form, float beta, const float* g_loss, int64_t n_elem_total,
                                 float* grad_weight, int64_t* hist_total, float* loss, int64_t* stats_total, void* stream) {
    if (int r = check_dims(0, K, D)) return r;
    if (!peer_bufs || world < 1 || world > VQ_PEER_MAX_RANKS || rank < 0 || rank >= world || (slot != 0 && slot != 1))
        return fail(VQ_ERR_ARG, "bad world / rank / slot for vq_backward_codebook_sharded");
    for (int r = 0; r < world; ++r)
        if (!peer_bufs[r]) return fail(VQ_ERR_ARG, "peer_bufs[%d] is NULL", r);
    if (!cb || !grad_weight || n_elem_total <= 0) return fail(VQ_ERR_ARG, "bad argument to vq_backward_codebook_sharded");
    if (form == VQ_FORM_VQGAN_L2) return fail(VQ_ERR_ARG, "VQ_FORM_VQGAN_L2 is not supported by the fused peer exchange (use the collective exchange)");
    vq::CodebookView cbv = vq::codebook_view(const_cast<void*>(cb), K, D);
    const float c2 = (form == VQ_FORM_VIT) ? 1.f : beta;
    const float coef = (float)((double)c2 * 2.0 / (double)n_elem_total);
    SlotTimer timer(static_cast<cudaStream_t>(stream), VQ_PROFILE_CODEBOOK_GRAD);
    VQ_CUDA(vq::launch_codebook_grad_sharded(peer_bufs, world, rank, slot, epoch, cbv, coef, g_loss, n_elem_total, form, beta,
                                             grad_weight, hist_total, loss, stats_total, nullptr, static_cast<cudaStream_t>(stream)));
    timer.stop();
    return VQ_OK;
}

int vq_backward_sharded(const void* const* peer_bufs, int world, int rank, int slot, uint32_t epoch, const float* g_zq, int64_t T,
                        const float* saved_zn, const float* saved_denom, const int64_t* idx, const void* cb, int K, int D, int form,
                        float beta, const float* g_loss, int64_t n_elem_total, float* grad_z, float* grad_weight,
                        int64_t* hist_total, float* loss, int64_t* stats_total, void* stream) {
    if (int r = check_dims(T, K, D)) return r;
    if (!peer_bufs || world < 1 || world > VQ_PEER_MAX_RANKS || rank < 0 || rank >= world || (slot != 0 && slot != 1))
        return fail(VQ_ERR_ARG, "bad world / rank / slot for vq_backward_sharded");
    for (int r = 0; r < world; ++r)
        if (!peer_bufs[r]) return fail(VQ_ERR_ARG, "peer_bufs[%d] is NULL", r);
    if (!cb || !grad_weight || !grad_z || !saved_zn || !saved_denom || !idx || n_elem_total <= 0)
        return fail(VQ_ERR_ARG, "bad argument to vq_backward_sharded");
    if (form == VQ_FORM_VQGAN_L2) return fail(VQ_ERR_ARG, "VQ_FORM_VQGAN_L2 is not supported by the fused peer exchange (use the collective exchange)");
    vq::CodebookView cbv = vq::codebook_view(const_cast<void*>(cb), K, D);
    const float c1 = (form == VQ_FORM_VIT) ? beta : 1.f, c2 = (form == VQ_FORM_VIT) ? 1.f : beta;
    const float coef1 = (float)((double)c1 * 2.0 / (double)n_elem_total), coef2 = (float)((double)c2 * 2.0 / (double)n_elem_total);
    vq::TokenBackward tok{g_zq, saved_zn, saved_denom, idx, T, coef1, grad_z};
    SlotTimer timer(static_cast<cudaStream_t>(stream), VQ_PROFILE_CODEBOOK_GRAD);
    VQ_CUDA(vq::launch_codebook_grad_sharded(peer_bufs, world, rank, slot, epoch, cbv, coef2, g_loss, n_elem_total, form, beta,
                                             grad_weight, hist_total, loss, stats_total, &tok, static_cast<cudaStream_t>(stream)));
    timer.stop();
    return VQ_OK;
}

int vq_gather_tokens(const void* idx, int token_bits, int64_t T, int64_t hw, const float* weight, const void* cb, int K, int D,
              int normalise, int layout_out, float* out, int64_t* stats, void* stream) {
    if (int r = check_dims(T, K, D)) return r;
    if (int r = check_layout(layout_out, T, hw)) return r;
    if (token_bits != 16 && token_bits != 32 && token_bits != 64) return fail(VQ_ERR_ARG, "token_bits %d: 16, 32 or 64", token_bits);
    if (!idx || !out) return fail(VQ_ERR_ARG, "idx/out is NULL");
    const float* table = nullptr;
    if (normalise) {
        if (!cb) return fail(VQ_ERR_ARG, "normalise=1 needs the prepared codebook");
        table = vq::codebook_view(const_cast<void*>(cb), K, D).en32;
    } else {
        if (!weight) return fail(VQ_ERR_ARG, "normalise=0 needs the raw weight");
        table = weight;
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (stats) VQ_CUDA(cudaMemsetAsync(stats + VQ_STAT_BAD_INDEX, 0, sizeof(int64_t), s));
    VQ_CUDA(vq::launch_gather(idx, T, hw, table, K, D, layout_out, out, stats, s, token_bits));
    return VQ_OK;
}

int vq_gather(const int64_t* idx, int64_t T, int64_t hw, const float* weight, const void* cb, int K, int D,
              int normalise, int layout_out, float* out, int64_t* stats, void* stream) {
    return vq_gather_tokens(idx, 64, T, hw, weight, cb, K, D, normalise, layout_out, out, stats, stream);
}

namespace {
int slot_total(int slot, double* ms_total, int64_t* launches) {
    double total = 0.0;
    auto& v = g_slot_events[slot];
    for (auto& p : v) {
        float ms = 0.f;
        VQ_CUDA(cudaEventSynchronize(p.second));
        VQ_CUDA(cudaEventElapsedTime(&ms, p.first, p.second));
        total += ms;
    }
    if (ms_total) *ms_total = total;
    if (launches) *launches = (int64_t)v.size();
    for (auto& p : v) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    v.clear();
    return VQ_OK;
}
}  // namespace

int vq_token_embed_tokens(const void* tokens, int token_bits, const uint8_t* mask, int64_t T, int64_t n_per_seq,
                          int64_t mask_token_id, int64_t ignore_index, const float* table, int64_t vocab, int dim,
                          const float* pos, float* embeds, int64_t* input_ids, int64_t* labels, int64_t* stats, void* stream) {
    if (T < 0 || T >= (1ll << 40)) return fail(VQ_ERR_ARG, "token count %lld out of range", (long long)T);
    if (token_bits != 16 && token_bits != 32 && token_bits != 64) return fail(VQ_ERR_ARG, "token_bits %d: 16, 32 or 64", token_bits);
    if (!tokens && T > 0) return fail(VQ_ERR_ARG, "tokens is NULL");
    if (embeds && (!table || vocab <= 0 || dim <= 0 || dim % 4 != 0))
        return fail(VQ_ERR_ARG, "embeds needs a (vocab, dim) table with dim a multiple of 4 (dim=%d)", dim);
    if (pos && n_per_seq <= 0) return fail(VQ_ERR_ARG, "pos needs n_per_seq > 0");
    if (!embeds && !input_ids && !labels) return fail(VQ_ERR_ARG, "nothing to produce");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (stats) VQ_CUDA(cudaMemsetAsync(stats + VQ_STAT_BAD_INDEX, 0, sizeof(int64_t), s));
    VQ_CUDA(vq::launch_token_embed(tokens, mask, T, n_per_seq > 0 ? n_per_seq : 1, mask_token_id, ignore_index, table, vocab, dim,
                                   pos, embeds, input_ids, labels, stats, s, nullptr, token_bits));
    return VQ_OK;
}

int vq_token_embed(const int64_t* tokens, const uint8_t* mask, int64_t T, int64_t n_per_seq, int64_t mask_token_id,
                   int64_t ignore_index, const float* table, int64_t vocab, int dim, const float* pos, float* embeds,
                   int64_t* input_ids, int64_t* labels, int64_t* stats, void* stream) {
    return vq_token_embed_tokens(tokens, 64, mask, T, n_per_seq, mask_token_id, ignore_index, table, vocab, dim, pos, embeds,
                                 input_ids, labels, stats, stream);
}

int vq_token_embed_causal_tokens(const void* tokens, int token_bits, int64_t T, int64_t n_per_seq, const float* table,
                                 int64_t vocab, int dim, const float* pos, const float* start, float* embeds, int64_t* stats,
                                 void* stream) {
    if (T < 0 || T >= (1ll << 40) || n_per_seq <= 0 || T % n_per_seq != 0)
        return fail(VQ_ERR_ARG, "vq_token_embed_causal needs T a multiple of n_per_seq > 0 (T=%lld)", (long long)T);
    if (token_bits != 16 && token_bits != 32 && token_bits != 64) return fail(VQ_ERR_ARG, "token_bits %d: 16, 32 or 64", token_bits);
    if (!tokens || !table || !start || !embeds || vocab <= 0 || dim <= 0 || dim % 4 != 0)
        return fail(VQ_ERR_ARG, "vq_token_embed_causal needs tokens, a (vocab, dim) table with dim %% 4 == 0, start and embeds");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (stats) VQ_CUDA(cudaMemsetAsync(stats + VQ_STAT_BAD_INDEX, 0, sizeof(int64_t), s));
    VQ_CUDA(vq::launch_token_embed(tokens, nullptr, T, n_per_seq, 0, 0, table, vocab, dim, pos, embeds, nullptr, nullptr, stats, s,
                                   start, token_bits));
    return VQ_OK;
}

int vq_token_embed_causal(const int64_t* tokens, int64_t T, int64_t n_per_seq, const float* table, int64_t vocab, int dim,
                          const float* pos, const float* start, float* embeds, int64_t* stats, void* stream) {
    return vq_token_embed_causal_tokens(tokens, 64, T, n_per_seq, table, vocab, dim, pos, start, embeds, stats, stream);
}

int vq_tokens_convert(const void* in, int in_bits, void* out, int out_bits, int64_t T, void* stream) {
    auto ok = [](int b) { return b == 16 || b == 32 || b == 64; };
    if (!ok(in_bits) || !ok(out_bits)) return fail(VQ_ERR_ARG, "token bits %d -> %d: 16, 32 or 64", in_bits, out_bits);
    if (T < 0 || ((!in || !out) && T > 0)) return fail(VQ_ERR_ARG, "bad argument to vq_tokens_convert");
    VQ_CUDA(vq::launch_tokens_convert(in, in_bits, out, out_bits, T, static_cast<cudaStream_t>(stream)));
    return VQ_OK;
}

int vq_embedding_backward_bytes(int64_t vocab, int dim, size_t* out) {
    if (!out || vocab <= 0 || dim <= 0) return fail(VQ_ERR_ARG, "bad argument to vq_embedding_backward_bytes");
    *out = vq::embedding_backward_bytes(vocab, dim);
    return VQ_OK;
}

int vq_embedding_backward(const int64_t* ids, int64_t n_ids, int64_t ids_per_seq, int64_t rows_per_seq, int64_t row_shift,
                          const float* grad_out, int64_t n_rows, int64_t hw, int64_t vocab, int dim, const void* cb,
                          float* grad_table, void* ws, size_t ws_bytes, void* stream) {
    if (n_ids < 0 || n_rows < 0 || vocab <= 0 || dim <= 0 || dim % 4 != 0)
        return fail(VQ_ERR_ARG, "bad sizes for vq_embedding_backward (dim must be a multiple of 4)");
    if ((n_ids > 0 && !ids) || (n_rows > 0 && !grad_out) || !grad_table) return fail(VQ_ERR_ARG, "ids / grad_out / grad_table is NULL");
    if (hw < 0 || (hw > 0 && n_ids % hw != 0)) return fail(VQ_ERR_ARG, "NCHW gradient needs hw dividing the id count");
    if (!ws || ws_bytes < vq::embedding_backward_bytes(vocab, dim)) return fail(VQ_ERR_WORKSPACE, "embedding backward workspace too small");
    vq::CodebookView cbv;
    if (cb) {
        if (int r = check_dims(0, (int)vocab, dim)) return r;
        cbv = vq::codebook_view(const_cast<void*>(cb), (int)vocab, dim);
    }
    VQ_CUDA(vq::launch_embedding_backward(ids, n_ids, ids_per_seq, rows_per_seq, row_shift, grad_out, n_rows, hw, vocab, dim,
                                          cb ? &cbv : nullptr, grad_table, ws, static_cast<cudaStream_t>(stream)));
    return VQ_OK;
}

int64_t vq_kernel_launches(void) { return (int64_t)vq::g_kernel_launches; }

int vq_profile_begin(int sample_every, unsigned slot_mask) {
    for (auto& v : g_slot_events) {
        for (auto& p : v) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
        v.clear();
    }
    g_launches_at_begin = vq::g_kernel_launches;
    g_profile_every = sample_every > 0 ? sample_every : 1;
    g_profile_mask = slot_mask ? slot_mask : ~0u;
    g_profile_calls = 0;
    g_sampled = false;
    g_profiling = true;
    return VQ_OK;
}

int vq_profile_end(double* search_ms_total, int64_t* search_launches, int64_t* kernel_launches) {
    g_profiling = false;
    g_sampled = false;
    if (kernel_launches) *kernel_launches = vq::g_kernel_launches - g_launches_at_begin;
    return slot_total(VQ_PROFILE_SEARCH, search_ms_total, search_launches);
}

int vq_profile_slot(int slot, double* ms_total, int64_t* launches) {
    if (slot < 0 || slot >= VQ_PROFILE_SLOTS) return fail(VQ_ERR_ARG, "profile slot %d out of range", slot);
    return slot_total(slot, ms_total, launches);
}

// ---- host-buffer step ---------------------------------------------------------------------------
namespace {
struct HostArena {
    void* cb; float* weight; float* z; float* g; float* zq; float* gz; int64_t* idx; float* zn; float* denom;
    int64_t* seg; float* gw; float* loss; int64_t* stats; int32_t* hist; void* fws; size_t fws_bytes; void* bws; size_t bws_bytes;
    size_t bytes;
};
HostArena carve_host(void* arena, int64_t T, int K, int D) {
    Bump b(arena);
    HostArena a;
    const size_t n = (size_t)(T > 0 ? T : 1);
    a.cb = b.take<char>(vq::codebook_bytes(K, D));
    a.weight = b.take<float>((size_t)K * D);
    a.z = b.take<float>(n * D);
    a.g = b.take<float>(n * D);
    a.zq = b.take<float>(n * D);
    a.gz = b.take<float>(n * D);
    a.idx = b.take<int64_t>(n);
    a.zn = b.take<float>(n * D);
    a.denom = b.take<float>(n);
    a.seg = b.take<int64_t>((size_t)K * D + K);
    a.gw = b.take<float>((size_t)K * D);
    a.loss = b.take<float>(64);
    a.stats = b.take<int64_t>(VQ_STATS_LEN);
    a.hist = b.take<int32_t>((size_t)K);
    a.fws_bytes = carve_forward(nullptr, T, K, D).bytes;
    a.fws = b.take<char>(a.fws_bytes);
    size_t bw = 0;
    vq_backward_workspace_bytes(T, K, D, &bw);
    a.bws_bytes = bw;
    a.bws = b.take<char>(bw);
    a.bytes = b.off;
    return a;
}
}  // namespace

int vq_host_step_arena_bytes(int64_t T, int K, int D, size_t* out) {
    if (!out) return fail(VQ_ERR_ARG, "out is NULL");
    if (int r = check_dims(T, K, D)) return r;
    *out = carve_host(nullptr, T, K, D).bytes;
    return VQ_OK;
}

namespace {
// copy streams and events of the chunked host step (created once per process, on first use)
constexpr int kMaxChunks = 64;
constexpr int64_t kChunkTokens = 32768;
struct HostPipe {
    cudaStream_t in = nullptr, out = nullptr;
    cudaEvent_t start = nullptr, weight = nullptr, tail = nullptr, finished = nullptr;
    cudaEvent_t loaded[kMaxChunks], done[kMaxChunks];
    bool ready = false;
};
// one pipe per device (streams and events belong to a device); the host-buffer step of one device is not re-entrant
// from several host threads at once -- it is a whole-batch call that owns the device for its duration
constexpr int kMaxDevices = 64;
HostPipe g_pipes[kMaxDevices];
HostPipe* current_pipe() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
    return &g_pipes[dev];
}
cudaError_t pipe_init(HostPipe& g_pipe) {
    if (g_pipe.ready) return cudaSuccess;
    cudaError_t e;
    if ((e = cudaStreamCreateWithFlags(&g_pipe.in, cudaStreamNonBlocking)) != cudaSuccess) return e;
    if ((e = cudaStreamCreateWithFlags(&g_pipe.out, cudaStreamNonBlocking)) != cudaSuccess) return e;
    cudaEvent_t* singles[4] = {&g_pipe.start, &g_pipe.weight, &g_pipe.tail, &g_pipe.finished};
    for (auto ev : singles)
        if ((e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming)) != cudaSuccess) return e;
    for (int i = 0; i < kMaxChunks; ++i) {
        if ((e = cudaEventCreateWithFlags(&g_pipe.loaded[i], cudaEventDisableTiming)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&g_pipe.done[i], cudaEventDisableTiming)) != cudaSuccess) return e;
    }
    g_pipe.ready = true;
    return cudaSuccess;
}
}  // namespace

int vq_host_step(const float* z_host, const float* g_zq_host, int64_t T, const float* weight_host, int K, int D,
                 int form, float beta, float* z_q_host, int64_t* idx_host, float* loss_host, float* grad_z_host,
                 float* grad_weight_host, int64_t* stats_host, void* dev_arena, size_t arena_bytes, void* stream) {
    if (int r = check_dims(T, K, D)) return r;
    if (!z_host || !weight_host || !z_q_host || !idx_host || !dev_arena) return fail(VQ_ERR_ARG, "NULL host/arena pointer");
    HostArena a = carve_host(dev_arena, T, K, D);
    if (arena_bytes < a.bytes) return fail(VQ_ERR_WORKSPACE, "arena too small: %zu < %zu", arena_bytes, a.bytes);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    HostPipe* pipe = current_pipe();
    if (!pipe) return fail(VQ_ERR_CUDA, "no current CUDA device for vq_host_step");
    VQ_CUDA(pipe_init(*pipe));
    HostPipe& p = *pipe;
    const size_t row_bytes = sizeof(float) * (size_t)D;
    const int64_t n_elem = T * D > 0 ? T * D : 1;
    const bool bwd = grad_z_host || grad_weight_host;
    // token chunks (multiples of 256 rows so that every chunk keeps the tensor-core path)
    int64_t chunk = T;
    if (T > 2 * kChunkTokens) {
        chunk = kChunkTokens;
        const int64_t need = (T + kMaxChunks - 1) / kMaxChunks;
        if (chunk < need) chunk = (need + 255) / 256 * 256;
    }
    const int n_chunks = T > 0 ? (int)((T + chunk - 1) / chunk) : 0;

    // ---- host -> device on the `in` stream, ordered after whatever `stream` was doing ----
    VQ_CUDA(cudaEventRecord(p.start, s));
    VQ_CUDA(cudaStreamWaitEvent(p.in, p.start, 0));
    VQ_CUDA(cudaStreamWaitEvent(p.out, p.start, 0));
    VQ_CUDA(cudaMemcpyAsync(a.weight, weight_host, row_bytes * K, cudaMemcpyHostToDevice, p.in));
    VQ_CUDA(cudaEventRecord(p.weight, p.in));
    for (int c = 0; c < n_chunks; ++c) {
        const int64_t t0 = c * chunk, n = (T - t0 < chunk) ? T - t0 : chunk;
        VQ_CUDA(cudaMemcpyAsync(a.z + t0 * D, z_host + t0 * D, row_bytes * n, cudaMemcpyHostToDevice, p.in));
        if (bwd && g_zq_host)
            VQ_CUDA(cudaMemcpyAsync(a.g + t0 * D, g_zq_host + t0 * D, row_bytes * n, cudaMemcpyHostToDevice, p.in));
        VQ_CUDA(cudaEventRecord(p.loaded[c], p.in));
    }
    // ---- kernels on `stream`, device -> host on the `out` stream ----
    VQ_CUDA(cudaStreamWaitEvent(s, p.weight, 0));
    if (int r = vq_codebook_prepare(a.weight, K, D, a.cb, vq::codebook_bytes(K, D), s)) return r;
    VQ_CUDA(cudaMemsetAsync(a.stats, 0, sizeof(int64_t) * VQ_STATS_LEN, s));
    VQ_CUDA(cudaMemsetAsync(a.hist, 0, sizeof(int32_t) * (size_t)K, s));
    if (grad_weight_host) VQ_CUDA(cudaMemsetAsync(a.seg, 0, sizeof(int64_t) * ((size_t)K * D + K), s));
    for (int c = 0; c < n_chunks; ++c) {
        const int64_t t0 = c * chunk, n = (T - t0 < chunk) ? T - t0 : chunk;
        VQ_CUDA(cudaStreamWaitEvent(s, p.loaded[c], 0));
        if (int r = vq_forward(a.z + t0 * D, VQ_LAYOUT_TOKEN_MAJOR, n, 0, nullptr, a.cb, K, D, form, beta, VQ_FLAG_KEEP_STATS, n_elem,
                               a.zq + t0 * D, a.idx + t0, nullptr, a.hist, a.stats, bwd ? a.zn + t0 * D : nullptr,
                               bwd ? a.denom + t0 : nullptr, grad_weight_host ? a.seg : nullptr, a.fws, a.fws_bytes, s))
            return r;
        if (grad_z_host)
            if (int r = vq_backward_tokens(g_zq_host ? a.g + t0 * D : nullptr, VQ_LAYOUT_TOKEN_MAJOR, n, 0, a.zn + t0 * D,
                                           a.denom + t0, a.idx + t0, nullptr, a.cb, K, D, form, beta, nullptr, n_elem,
                                           a.gz + t0 * D, nullptr, a.bws, a.bws_bytes, s))
                return r;
        VQ_CUDA(cudaEventRecord(p.done[c], s));
        VQ_CUDA(cudaStreamWaitEvent(p.out, p.done[c], 0));
        VQ_CUDA(cudaMemcpyAsync(z_q_host + t0 * D, a.zq + t0 * D, row_bytes * n, cudaMemcpyDeviceToHost, p.out));
        VQ_CUDA(cudaMemcpyAsync(idx_host + t0, a.idx + t0, sizeof(int64_t) * n, cudaMemcpyDeviceToHost, p.out));
        if (grad_z_host)
            VQ_CUDA(cudaMemcpyAsync(grad_z_host + t0 * D, a.gz + t0 * D, row_bytes * n, cudaMemcpyDeviceToHost, p.out));
    }
    // ---- whole-batch tail: loss and codebook gradient from the accumulated fixed-point sums ----
    if (grad_weight_host) {   // the chunks' forwards accumulated the integer segment sums of the whole batch
        if (int r = vq_backward_codebook(a.seg, a.cb, K, D, form, beta, nullptr, n_elem, a.gw, a.stats, loss_host ? a.loss : nullptr, s))
            return r;
    } else if (loss_host) {
        VQ_CUDA(vq::launch_loss_finalize(a.stats, n_elem, form, beta, a.loss, s));
    }
    VQ_CUDA(cudaEventRecord(p.tail, s));
    VQ_CUDA(cudaStreamWaitEvent(p.out, p.tail, 0));
    if (loss_host) VQ_CUDA(cudaMemcpyAsync(loss_host, a.loss, sizeof(float), cudaMemcpyDeviceToHost, p.out));
    if (grad_weight_host) VQ_CUDA(cudaMemcpyAsync(grad_weight_host, a.gw, row_bytes * K, cudaMemcpyDeviceToHost, p.out));
    if (stats_host) VQ_CUDA(cudaMemcpyAsync(stats_host, a.stats, sizeof(int64_t) * VQ_STATS_LEN, cudaMemcpyDeviceToHost, p.out));
    VQ_CUDA(cudaEventRecord(p.finished, p.out));
    VQ_CUDA(cudaStreamWaitEvent(s, p.finished, 0));
    return VQ_OK;
}

}  // extern "C"
