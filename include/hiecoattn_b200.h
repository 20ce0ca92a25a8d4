/*
 * hiecoattn_b200 -- C ABI of the B200-native Hierarchical Co-Attention hot path.
 *
 * The reference (Axe--/Visual-Question-Answering) has no native code: its hot path is the ATen calls
 * issued by four nn.Modules in model.py.  This header is therefore the boundary a maintainer of the
 * reference would bind (ctypes stub in INTEGRATION.md); each entry point names the reference code it
 * replaces.  The Python package visual-question-answering_b200 wraps these as torch.library custom
 * ops (namespace "hiecoattn") behind modules with the reference's names and signatures.
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer on the current CUDA device
 *     unless stated otherwise; all matrices are dense row-major fp32 unless stated otherwise.
 *   - `stream` is a cudaStream_t passed as void*; nothing synchronises the host, nothing runs on the
 *     legacy default stream unless that is the stream passed in; safe under CUDA-graph capture.
 *   - the library allocates nothing per call: outputs, saved tensors and scratch (`ws`, sized by the
 *     matching *_workspace query, 256-byte aligned) are owned by the caller.
 *   - return value 0 = ok; non-zero = error, message via hca_last_error() (thread-local).  Never
 *     throws, never exits.
 *   - there is NO CPU implementation behind this ABI.
 */
#ifndef HIECOATTN_B200_H_
#define HIECOATTN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HCA_ABI_VERSION 6

#if defined(__GNUC__)
#define HCA_API __attribute__((visibility("default")))
#else
#define HCA_API
#endif

/* error codes */
#define HCA_OK 0
#define HCA_ERR_ARG 1      /* bad shape / null pointer / unsupported size */
#define HCA_ERR_CUDA 2     /* CUDA runtime / driver error */
#define HCA_ERR_WORKSPACE 3 /* workspace too small */

HCA_API int hca_abi_version(void);
HCA_API const char* hca_last_error(void);
/* number of CUDA kernels this library has launched from the calling process so far (bench.py's
 * gpu_launches is the difference across the timed region) */
HCA_API int64_t hca_launch_count(void);
/* runtime switches.  "pdl" = "0" / "1": programmatic dependent launch off / on (a profiler leg that wants exclusive kernel
 * durations turns it off); "pool_tie_cap" = "<n>": test hook, caps the near-tie list of the phrase max-pool repair at n entries
 * (0 = default) so that its exhaustive mode can be exercised; "gemm" reports "tc" (tcgen05: the only contraction backend).
 * Returns 0 or HCA_ERR_ARG. */
HCA_API int hca_set_option(const char* name, const char* value);
HCA_API const char* hca_get_option(const char* name);

/* page-locked host staging buffers for the loader -> GPU hand-over of a step's inputs (replaces the pageable .to(device) copies of
 * main.py:205-208); write_combined != 0: write-combined pages (CPU-write-only staging, cheaper for the device to read) */
HCA_API int hca_pinned_alloc(size_t bytes, int write_combined, void** out);
HCA_API int hca_pinned_free(void* p);

/* ---- question encoder: embedding gather (replaces nn.Embedding, model.py:263,282) -------------- */
/* out[r,:] = table[tokens[r],:] for r < rows; tokens int64 in [0,vocab).  E % 4 == 0. */
HCA_API int hca_embedding_fwd(const int64_t* tokens, const float* table, float* out,
                      int64_t rows, int E, int64_t vocab, void* stream);
/* dtable (pre-zeroed by this call) += scatter of dout rows; row 0 (padding_idx) stays zero. */
HCA_API int hca_embedding_bwd(const int64_t* tokens, const float* dout, float* dtable,
                      int64_t rows, int E, int64_t vocab, void* stream);

/* ---- PhraseConvPool (replaces model.py:304-334) ------------------------------------------------ */
/* x [B,T,E]; w{1,2,3} = Conv1d weights [E,E,k] (k=1,2,3), b{1,2,3} [E];
 * lens: optional int64 [B] (device); when non-null rows t >= lens[b] of `out` are zeroed (the
 * pack/pad step of model.py:287-292) and idx there is 0.
 * out [B,T,E]; idx [B,T,E] uint8 = position (0..2) inside the consecutive channel triple of
 * [uni|bi|tri] that attained the max, first index on ties (MaxPool2d((1,3)), model.py:311,329-332). */
HCA_API size_t hca_phrase_conv_pool_workspace(int B, int T, int E);
/* `fsaved` (optional, hca_phrase_conv_pool_saved_bytes, 256-byte aligned, opaque): when non-null the forward leaves the bf16 operand
 * planes of the row-shifted input and of the three weights there, and a backward call that is handed the same buffer reuses them instead
 * of converting the operands again (the weights do not change between the forward and the backward of a step). */
HCA_API size_t hca_phrase_conv_pool_saved_bytes(int B, int T, int E);
HCA_API int hca_phrase_conv_pool_fwd(const float* x, const float* w1, const float* b1, const float* w2,
                             const float* b2, const float* w3, const float* b3, const int64_t* lens,
                             float* out, uint8_t* idx, void* fsaved, size_t fsaved_bytes, int B, int T, int E,
                             void* ws, size_t ws_bytes, void* stream);
/* gradients of the above given dout [B,T,E]; dx may be null (input does not need grad); fsaved may be null. */
HCA_API int hca_phrase_conv_pool_bwd(const float* x, const float* w1, const float* w2, const float* w3,
                             const float* out, const uint8_t* idx, const float* dout, const int64_t* lens,
                             const void* fsaved, size_t fsaved_bytes,
                             float* dx, float* dw1, float* db1, float* dw2, float* db2, float* dw3, float* db3,
                             int B, int T, int E, void* ws, size_t ws_bytes, void* stream);

/* ---- sentence LSTM (replaces pack_padded_sequence -> nn.LSTM(E,H) -> pad_packed_sequence, model.py:269,287-296) ---- */
/* x [B,T,E]; lens int64 [B] (device), each sequence runs for lens[b] steps from h0 = c0 = 0; w_ih [4H,E], w_hh [4H,H],
 * b_ih, b_hh [4H] in PyTorch's layout (gates i, f, g, o stacked along rows).  out [B,T,H], rows t >= lens[b] zero.
 * Supported: E % 8 == 0, H % 16 == 0, H <= 512 (hca_lstm_supported returns 1); `saved` = one opaque 256-byte aligned
 * buffer of hca_lstm_saved_bytes bytes (gate activations, cell states, bf16 hi/lo planes of x and of the hidden states). */
HCA_API int hca_lstm_supported(int B, int T, int E, int H);
HCA_API size_t hca_lstm_saved_bytes(int B, int T, int E, int H);
HCA_API size_t hca_lstm_workspace(int B, int T, int E, int H);
HCA_API int hca_lstm_fwd(const float* x, const int64_t* lens, const float* w_ih, const float* w_hh,
                 const float* b_ih, const float* b_hh, float* out, void* saved, size_t saved_bytes,
                 int B, int T, int E, int H, void* ws, size_t ws_bytes, void* stream);
/* dout [B,T,H] -> dx [B,T,E] (may be null), dw_ih [4H,E], dw_hh [4H,H], db_ih, db_hh [4H]; gradients of positions
 * t >= lens[b] are ignored (their outputs are the constant zero). */
HCA_API int hca_lstm_bwd(const int64_t* lens, const float* w_ih, const float* w_hh, const void* saved, size_t saved_bytes,
                 const float* dout, float* dx, float* dw_ih, float* dw_hh, float* db_ih, float* db_hh,
                 int B, int T, int E, int H, void* ws, size_t ws_bytes, void* stream);

/* ---- ParallelCoAttention, all three levels (replaces model.py:356-397) -------------------------- */
/* V [B,N,d] with ELEMENT strides (v_sb, v_sn, v_sd) -- the reference hands a permuted VGG view
 * (model.py:217); q0,q1,q2 = word / phrase / sentence features, each dense [B,T,d];
 * Wv,Wq [d,d] (nn.Linear layout [out,in]), bv,bq [d], wv,wq [d], cv,cq device scalars [1].  d % 8 == 0.
 * W_b is declared by the reference (model.py:347) but never used (model.py:377): it is not an input.
 * outputs: vhat [3,B,d], qhat [3,B,d];
 * saved for backward: one opaque, 256-byte aligned buffer of hca_coattn_saved_bytes(B,N,T,d) bytes (bf16 hi/lo
 * operand planes of V, the stacked question levels, PV, PQ, C and the attention weights a_v, a_q). */
HCA_API size_t hca_coattn_saved_bytes(int B, int N, int T, int d);
HCA_API size_t hca_coattn_workspace(int B, int N, int T, int d, int need_dv);
HCA_API int hca_coattn_fwd(const float* V, int64_t v_sb, int64_t v_sn, int64_t v_sd,
                   const float* q0, const float* q1, const float* q2,
                   const float* Wv, const float* bv, const float* Wq, const float* bq,
                   const float* wv, const float* cv, const float* wq, const float* cq,
                   float* vhat, float* qhat, void* saved, size_t saved_bytes,
                   int B, int N, int T, int d, void* ws, size_t ws_bytes, void* stream);
/* Byte offsets, inside the `saved` buffer hca_coattn_fwd has filled, of the attention weights (fp32): a_v [B][3][N] and
 * a_q [B][3][T] (level order word, phrase, sentence; softmax over regions / over all T token positions, model.py:387-388).
 * For attention-map export at inference time (README "Inference", TO-DO in the reference). */
HCA_API int hca_coattn_saved_attention(int B, int N, int T, int d, size_t* av_offset, size_t* aq_offset);
/* gvhat,gqhat [3,B,d] = dL/dvhat, dL/dqhat.  Outputs: dQ [3,B,T,d] (dQ[l] = gradient of q_l); dWv,dWq [d,d];
 * dbv,dbq,dwv,dwq [d]; dcv,dcq [1]; dV [B,N,d] dense or null when the image features need no gradient (frozen
 * VGG, main.py:67).  Weight gradients are summed over batch and levels. */
HCA_API int hca_coattn_bwd(const float* Wv, const float* Wq, const float* wv, const float* wq,
                   const void* saved, size_t saved_bytes, const float* gvhat, const float* gqhat,
                   float* dV, float* dQ, float* dWv, float* dbv, float* dWq, float* dbq,
                   float* dwv, float* dcv, float* dwq, float* dcq,
                   int B, int N, int T, int d, void* ws, size_t ws_bytes, void* stream);

/* ---- MLPClassifier (replaces model.py:414-434) -------------------------------------------------- */
/* vhat,qhat [3,B,d] (levels word, phrase, sentence); Ww [d,d], Wp [d,2d], Ws [mlp,2d], Wh [K,mlp]; d % 8 == 0.
 * outputs: logits [B,K]; `saved` (hca_mlp_saved_bytes, 256-byte aligned, opaque): bf16 hi/lo operand planes of the four
 * weights, of xw [B,d], xp [B,2d] (= [q_p+v_p | h_w]), xs [B,2d] (= [q_s+v_s | h_p]) and of hs [B,mlp]. */
HCA_API size_t hca_mlp_workspace(int B, int d, int mlp, int K);
HCA_API size_t hca_mlp_saved_bytes(int B, int d, int mlp, int K);
HCA_API int hca_mlp_fwd(const float* vhat, const float* qhat,
                const float* Ww, const float* bw, const float* Wp, const float* bp,
                const float* Ws, const float* bs, const float* Wh, const float* bh,
                float* logits, void* saved, size_t saved_bytes,
                int B, int d, int mlp, int K, void* ws, size_t ws_bytes, void* stream);
/* dlogits [B,K] -> g [3,B,d] (gradient of BOTH vhat and qhat of each level: they enter as q+v),
 * and all weight / bias gradients (written, not accumulated). */
HCA_API int hca_mlp_bwd(const float* dlogits, const void* saved, size_t saved_bytes,
                float* g, float* dWw, float* dbw, float* dWp, float* dbp, float* dWs, float* dbs,
                float* dWh, float* dbh, int B, int d, int mlp, int K,
                void* ws, size_t ws_bytes, void* stream);

/* ---- optimizer step (replaces torch.optim.Adam(model.parameters(), lr), main.py:180,222) ------------------------ */
/* One fused pass over flat fp32 buffers p, g, m, v of n elements (n % 4 == 0, 16-byte aligned): Adam with bias correction,
 * no weight decay, no amsgrad.  `step` (device int64[1]) is advanced by one and `coef` (device float[2], scratch) receives
 * the bias-correction factors, so the call can be captured in a CUDA graph. */
HCA_API int hca_adam_step(float* p, const float* g, float* m, float* v, int64_t n, long long* step, float* coef,
                  float lr, float beta1, float beta2, float eps, void* stream);

/* advances `step` and fills `coef` only (the first half of hca_adam_step), for callers that apply the update with
 * hca_dp_reduce_adam below */
HCA_API int hca_adam_prep(long long* step, float* coef, float lr, float beta1, float beta2, void* stream);

/* ---- data parallel: gradient all-reduce + Adam + parameter broadcast in ONE kernel over NVLink / NVSwitch ----------
 * (the reference has no multi-GPU path: main.py:102-106 is a commented-out nn.DataParallel TODO; this is the collective
 * `north_star` adds, fused with the optimizer of main.py:180,222.)
 * Every rank holds one SYMMETRIC block (same layout on every GPU, mapped into every peer): [flags | ... g ... | ... p ...].
 *   peer_bases  HOST array [world] of the block's base address on every rank as mapped into THIS process (peer_bases[rank] is
 *               the local one); mc_base = the block's multicast (NVLS) address, or 0 to use plain peer loads / stores;
 *   flags_off   byte offset of hca_dp_flags_bytes() bytes of zero-initialised flag words; g_off / p_off: byte offsets of the
 *               flat fp32 gradient / parameter buffers; [begin, end) the element range to process (multiples of 4);
 *   channel     0..3: launches that may run concurrently (different streams) must use different channels;
 *   max_ctas    cap on the grid (0 = no cap): a launch meant to overlap other work leaves SMs to it;
 *   mode        bit 0: Adam update of the owner's slice + broadcast of the new parameters to every rank (m, v: local fp32 [n]
 *               moment buffers, of which each rank uses only its slice; coef from hca_adam_prep);
 *               bit 1: the summed gradient is written back to every rank's g (a plain all-reduce).
 * Collective: every rank must make the same call (same range, channel, max_ctas, mode).  A rank that never arrives trips a
 * bounded wait (the kernel traps, the host sees a CUDA error): no hang. */
HCA_API size_t hca_dp_flags_bytes(void);
HCA_API int hca_dp_reduce_adam(const uint64_t* peer_bases, uint64_t mc_base, size_t flags_off, size_t g_off, size_t p_off,
                       int64_t begin, int64_t end, int rank, int world, int channel, int max_ctas, float* m, float* v,
                       const float* coef, float beta1, float beta2, float eps, int mode, void* stream);

/* ---- loss (replaces nn.CrossEntropyLoss()(logits, labels) and its backward, main.py:179,214) ---------------------- */
/* loss[0] = scale * mean_b CE(logits[b,:K], labels[b]); dlogits (optional, leading dim ldd) = d loss / d logits.  One launch.
 * `ws` >= hca_ce_loss_workspace(B) bytes of scratch.  labels int64 in [0,K): others give NaN. */
HCA_API size_t hca_ce_loss_workspace(int B);
HCA_API int hca_ce_loss(const float* logits, int64_t ld, const int64_t* labels, int B, int K, float scale, float* loss,
                float* dlogits, int64_t ldd, void* ws, size_t ws_bytes, void* stream);

/* ---- building block exposed for tests and profiling: one dense contraction ------------------------- */
/* fp32 row-major in and out.  layout 0 "nt": D[M,N] = A[M,K] . B[N,K]^T (+bias[N])   (nn.Linear forward)
 *                            layout 1 "nn": D[M,N] = A[M,K] . B[K,N]     (+bias[N])   (data gradient)
 *                            layout 2 "tn": D[M,N] = A[K,M]^T . B[K,N]                (weight gradient, split-K)
 * path 0 = fp32 CUDA cores; 1 = tcgen05 with bf16x2 operand splitting (3 MMAs, ~2^-16 operand
 * precision); 2 = tcgen05 bf16x3 (6 MMAs, fp32-grade). */
HCA_API size_t hca_gemm_workspace(int M, int N, int K);
/* The projection kernel of the co-attention (PV = V.Wv^T + bv, model.py:380-384) on its own, operands already in the library's
 * operand format: bf16 hi/lo planes [2][rows][cols] (x = hi + lo).  hca_split_planes produces that format from fp32 [rows, cols];
 * hca_proj_planes computes out planes [2][M][N] = A[M,K] . W[N,K]^T + bias[N] with one tcgen05 GEMM launch (3 MMAs per product). */
HCA_API int hca_split_planes(const float* src, int64_t rows, int cols, void* planes, void* stream);
HCA_API int hca_proj_planes(const void* a_planes, int64_t M, int K, const void* w_planes, int N, const float* bias,
                    void* out_planes, void* stream);
/* debug / profiling aid: subsequent tensor-core GEMM launches record per-CTA clock64() stamps into
 * buf [nctas][64] int64 ([0..7]: start, setup done, first tile landed, MMAs issued, epilogue start,
 * epilogue end, CTA end, SM id; [8+i], [24+i], [40+i]: per-k-block producer / landed / issued stamps); pass NULL to switch it off. */
HCA_API int hca_debug_gemm_timeline(void* buf, int nctas);
/* same, but only the launch_index-th tensor-core GEMM launch after this call records ([8..63] then hold the epilogue
 * stamps of one non-leader epilogue thread: per 32-column chunk {buffer free, addend landed, operands in registers,
 * tile staged, group barrier passed}) */
HCA_API int hca_debug_gemm_timeline_select(void* buf, int nctas, int launch_index);
/* debug: CTA 0 of the following LSTM recurrence launches records clock64() stamps into buf (int64 [7 rounds][8]); NULL = off */
HCA_API int hca_debug_lstm_timeline(void* buf);
/* profiling aid (bench.py's roofline legs): the recurrence kernel launched by the following hca_lstm_fwd (which = 0) /
 * hca_lstm_bwd (which = 1) calls is bracketed by the two cudaEvent_t handles on the launching stream; NULLs switch it off */
HCA_API int hca_debug_lstm_events(void* ev_start, void* ev_stop, int which);
/* one weight-gradient product on its own: D[M,N] += A[K,M]^T . B[K,N] from bf16 hi/lo planes [2][K][M], [2][K][N] (split-K,
 * fp32 reduce-add into D) */
HCA_API int hca_wgrad_planes(const void* a_planes, const void* b_planes, int M, int N, int64_t K, float* D, void* stream);
/* one dense contraction from fp32 operands (tests / profiling): layout 0 nt, 1 nn, 2 tn; path 1 = bf16x2 split (3 MMAs), 2 = bf16x3 */
HCA_API int hca_gemm(const float* A, const float* B, const float* bias, float* D, int M, int N, int K,
                     int layout, int path, void* ws, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HIECOATTN_B200_H_ */
