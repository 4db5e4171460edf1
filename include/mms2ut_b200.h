/* C ABI of libmms2ut_b200.so -- the sm_100a kernels behind the mm_s2ut_transformer encoder hot path.
 *
 * Every entry point is `extern "C"`, takes plain device pointers + sizes + a CUDA stream handle
 * (cudaStream_t passed as void*), launches asynchronously on that stream, never allocates or frees
 * device memory (the caller owns every buffer incl. workspaces) and returns a cudaError_t-compatible
 * int (0 = success; argument errors return cudaErrorInvalidValue = 1).  Re-entrant; no global state
 * except the lazily resolved driver entry point for tensor-map encoding.
 *
 * The reference (whxhcj/multimodal-S2UT) is pure Python on fairseq/torchaudio; each group below names
 * the reference call site it replaces (paths relative to the reference root).
 */
#ifndef MMS2UT_B200_H_
#define MMS2UT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MM_DTYPE_BF16 0
#define MM_DTYPE_F16 1

/* Library self-description: returns the ABI version (checked by the Python loader). */
int mm_abi_version(void);
/* Name of the last failing CUDA call inside the library on this thread ("" if none). */
const char* mm_last_error(void);

/* ---------------------------------------------------------------------------------------------
 * Front-end.  Replaces mm_s2ut/data/audio_utils.py:326-349 (get_fbank -> fairseq _get_torchaudio_fbank
 * -> torchaudio.compliance.kaldi.fbank(num_mel_bins=80)) and the UtteranceCMVN feature transform
 * applied at mm_s2ut/data/speech_to_speech_dataset.py:271-273, plus fairseq _collate_frames' zero padding.
 *
 * mm_fbank_f32: wav [B, wav_stride] fp32 (already x 2^15), n_samples [B] int64 ->
 *   feats [B, max_frames, 80] fp32 raw log-mel (rows >= the utterance's frame count are NOT written).
 * mm_cmvn_stats: per-utterance, per-bin (mean, std) -> mean_std [B, 2, 80] fp32, computed with the reference's
 *   own arithmetic (numpy fp32, sequential accumulation over frames, var = E[x^2] - mean^2 floored at 1e-10),
 *   so that its rounding behaviour -- part of the reference's result -- is reproduced.
 * mm_cmvn_apply: y = (x - mean) / std and zero padding; writes either/both of
 *   out_f32 [B, max_frames, 80] and out_op [B, op_frames, 80] (16-bit operand type; row r of the utterance
 *   lands at row r + op_row_offset; all other rows are written as 0).  With mean_std == NULL the input is
 *   taken as already normalised features (the reference's 3-D src_tokens contract) and only copied/padded.
 * --------------------------------------------------------------------------------------------- */
int mm_fbank_f32(const float* wav, const int64_t* n_samples, int32_t batch, int64_t wav_stride, float* feats,
                 int32_t max_frames, const float* tables, void* stream);
/* Same, from raw int16 PCM (what the wav files hold before mm_s2ut/data/audio_utils.py:281-290 turns them into
 * float32 / 2^15 and multiplies by 2^15 again): the conversion is exact, the dominant read is halved. */
int mm_fbank_i16(const int16_t* wav, const int64_t* n_samples, int32_t batch, int64_t wav_stride, float* feats,
                 int32_t max_frames, const float* tables, void* stream);
/* Constant tables of the fbank kernel (povey window, FFT twiddles, sparse mel bank): the caller allocates
 * mm_fbank_table_floats() floats on the host, fills them with mm_fbank_build_tables() and keeps a device copy
 * that it passes to mm_fbank_f32 (the library itself owns no device memory). */
int mm_fbank_table_floats(void);
int mm_fbank_build_tables(float* host_out);
int mm_cmvn_stats(const float* feats, const int64_t* n_samples_or_frames, int32_t lengths_are_samples,
                  int32_t batch, int32_t max_frames, float* mean_std, void* stream);
int mm_cmvn_apply(const float* feats, const float* mean_std, const int64_t* n_samples_or_frames,
                  int32_t lengths_are_samples, int32_t batch, int32_t max_frames, float* out_f32, void* out_op,
                  int32_t op_frames, int32_t op_row_offset, int32_t dtype, void* stream);
/* mm_cmvn_apply with the train-time SpecAugment of the reference's data config (`_train: [utterance_cmvn, specaugment]`,
 * fairseq SpecAugmentTransform applied at mm_s2ut/data/speech_to_speech_dataset.py:271-273) fused into the same pass:
 * spec_masks [B][2 * (n_fmask + n_tmask)] int32 = n_fmask (f0, f) frequency bands then n_tmask (t0, t) frame ranges per
 * utterance, drawn on the host with the reference's numpy calls; masked cells are written as mask_value. */
int mm_cmvn_apply_specaug(const float* feats, const float* mean_std, const int64_t* n_samples_or_frames,
                          int32_t lengths_are_samples, int32_t batch, int32_t max_frames, float* out_f32, void* out_op,
                          int32_t op_frames, int32_t op_row_offset, int32_t dtype, const int32_t* spec_masks,
                          int32_t n_fmask, int32_t n_tmask, float mask_value, void* stream);
/* out_lens[b] = conv-subsampled length of utterance b (int32): frames -> floor((L-1)/2+1) n_layers times;
 * frames = 1 + (n - 400) / 160 when lengths_are_samples.  (fairseq Conv1dSubsampler.get_out_seq_lens_tensor) */
int mm_seq_lens(const int64_t* n_samples_or_frames, int32_t lengths_are_samples, int32_t batch, int32_t n_layers,
                int32_t* out_lens, void* stream);
/* mm_seq_lens and mm_padding_mask (below) in one launch: out_lens [batch] and mask [batch, T]. */
int mm_seq_lens_mask(const int64_t* n_samples_or_frames, int32_t lengths_are_samples, int32_t batch, int32_t n_layers,
                     int32_t T, int32_t* out_lens, uint8_t* mask, void* stream);
/* mask[b, t] = (t >= seq_lens[b]) as bytes (PyTorch bool layout): fairseq lengths_to_padding_mask applied to the
 * subsampled lengths, i.e. the encoder_padding_mask of the encoder-out dict (mm_s2s_transformer.py:378-562). */
int mm_padding_mask(const int32_t* seq_lens, int32_t batch, int32_t T, uint8_t* mask, void* stream);

/* ---------------------------------------------------------------------------------------------
 * tcgen05 GEMM with fused epilogues:  acc[r, c] = sum_k A[r, k] * W[c, k]   (fp32 accumulation in TMEM)
 * A is 16-bit [batches][rows][K] (strides in elements; a row stride smaller than K expresses the
 * overlapping windows of a stride-2 Conv1d over a time-major buffer), optionally split along K into two
 * tensors (a0: first k_split columns, a1: the rest).  W is 16-bit [N][K] (PyTorch Linear layout), optionally
 * batched.  Replaces, by epilogue mode:
 *   MM_EPI_OP        q/k/v projections (fairseq MultiheadAttention; fuse.py:76-80) -> 16-bit out, optional
 *                    column scale (q * head_dim^-0.5), columns >= vt_col0 stored transposed (V^T) for attention
 *   MM_EPI_RELU_OP   fc1 + ReLU (fairseq TransformerEncoderLayer)
 *   MM_EPI_RESID_F32 out_proj / fc2 + bias + residual on the fp32 residual stream
 *   MM_EPI_GLU_OP    Conv1d(k5,s2)+GLU #1 of fairseq Conv1dSubsampler (reached at mm_s2s_transformer.py:464)
 *   MM_EPI_GLU_POS_F32  Conv1d+GLU #2, x*sqrt(d) + sinusoidal position (S2TTransformerEncoder._forward)
 *   MM_EPI_F32_OP    SelectiveAttention.proj (fuse.py:115-116): fp32 and 16-bit copies
 *   MM_EPI_GATE      selective gate, mm_s2s_transformer.py:612-618: g = sigmoid(acc + b),
 *                    out = (1-g)*text + g*attn, stored as T x B x C
 *   MM_EPI_F32       plain fp32 store (attention scores q k^T, fuse.py:87)
 *   MM_EPI_MASK_OP   backward of fc1's ReLU (+ activation dropout) fused into the fc2 dgrad (autograd of fairseq
 *                    TransformerEncoderLayer): out = aux0 > 0 ? (acc + bias) * scale : 0, 16-bit, aux0 = the kept
 *                    16-bit activation [rows, n]
 * --------------------------------------------------------------------------------------------- */
enum {
  MM_EPI_OP = 0,
  MM_EPI_RELU_OP = 1,
  MM_EPI_RESID_F32 = 2,
  MM_EPI_GLU_OP = 3,
  MM_EPI_GLU_POS_F32 = 4,
  MM_EPI_F32_OP = 5,
  MM_EPI_GATE = 6,
  MM_EPI_F32 = 7,
  MM_EPI_MASK_OP = 8
};

typedef struct mm_gemm_args {
  const void* a0;      /* [batches][rows][k_split] */
  const void* a1;      /* [batches][rows][k - k_split] or NULL */
  const void* w;       /* [w_batches][n][k] */
  int64_t a0_ld, a0_bs, a1_ld, a1_bs, w_ld, w_bs; /* row / batch strides, elements */
  int32_t rows, batches, n, k, k_split, w_batched;
  int32_t dtype;       /* MM_DTYPE_* of a0/a1/w and of every "op" output */
  int32_t mode;        /* MM_EPI_* */
  int32_t block_n;     /* 128 or 256; 0 = library default */
  const float* bias;   /* [n] or NULL */
  float scale;         /* MM_EPI_OP: columns < scale_cols multiplied by scale; GLU_POS: embed scale */
  int32_t scale_cols;
  void* out0;          /* primary output */
  int64_t out0_ld, out0_bs;
  void* out1;          /* secondary output (MM_EPI_F32_OP: the 16-bit copy) */
  int64_t out1_ld, out1_bs;
  const void* aux0;    /* RESID: residual ; GATE: text (fp32) ; MASK_OP: kept activation (16-bit, row stride aux_ld) */
  const float* aux1;   /* GATE: attention output (fp32) */
  int64_t aux_ld;
  int32_t rows_per_seq; /* >0: batches==1 and row r is (b, t) = divmod(r, rows_per_seq) */
  int32_t out_tbc;      /* RESID/GATE: store row (b,t) at (t*n_seqs + b) */
  int32_t n_seqs;
  int32_t out_row_offset; /* GLU_OP: output row t lands at t + out_row_offset (leading zero frames) */
  void* vt;             /* MM_EPI_OP: transposed output [n_seqs][vt_rows][vt_ld] for columns >= vt_col0 */
  int32_t vt_col0, vt_rows;
  int64_t vt_ld;
  const float* pos;     /* GLU_POS: [>= rows+2][n/2] sinusoidal table */
  const int32_t* seq_lens; /* GLU_POS: valid length per sequence */
  /* Operand layouts of the backward pass (all 0 in the forward pass):
   *   a_mn / w_mn     the operand is MN-major: memory is [contraction index][row], ld = elements between contraction
   *                   indices.  dgrad (dX = dY W) reads W [n_out, k_in] as stored; wgrad (dW = dY^T X) reads dY and X as
   *                   stored -- no transposed copies.  Needs k % 64 == 0 unless the operand ends at k.
   *   a_kbatch / w_kbatch  (MN-major only) the batch index advances the contraction index by k (split-K over tokens):
   *                   batch b covers contraction rows [b k, (b+1) k) of a tensor with *_k_total rows (zero beyond).
   *   heads, head_stride, a_hm / w_hm / out_hm   batch index = (sequence, head): the operand / output is
   *                   [sequence][rows][heads * head_stride] and head h uses the column block starting at h * head_stride
   *                   (attention backward straight from / into the q|k|v layout).  out_hm: MM_EPI_OP, n % 64 == 0. */
  int32_t a_mn, w_mn, a_kbatch, w_kbatch, a_hm, w_hm, out_hm, heads, head_stride;
  int64_t a_k_total, w_k_total;
  /* MM_EPI_RELU_OP, training forward: activation dropout on relu(acc + bias) (fairseq --activation-dropout /
   * --relu-dropout, scripts/textless/1_train.sh:112): kept values are scaled by 1 / (1 - drop_p); the mask of element
   * (row, column) is the counter-based one of mm_dropout at index row * n + column, site drop_site, seed drop_seed +
   * *drop_seed_dev (drop_seed_dev optional).  drop_p == 0: off. */
  float drop_p;
  uint32_t drop_site;
  uint64_t drop_seed;
  const uint64_t* drop_seed_dev;
} mm_gemm_args;

int mm_gemm(const mm_gemm_args* args, void* stream);

/* Fused sub-layer tail of the pre-LN fairseq TransformerEncoderLayer (out_proj or fc2) plus the LayerNorm that
 * opens the next sub-layer (self_attn_layer_norm / final_layer_norm / the encoder's last layer_norm):
 *     x[r, :] <- x[r, :] + a[r, :] W^T + bias          (fp32 residual stream, updated in place)
 *     h[r, :] <- LayerNorm(x[r, :]) * gamma + beta      (16-bit operand copy, optional fp32 copy)
 * a [rows, k] and w [n, k] are 16-bit, n must be 512 (a full row lives in one TMEM accumulator). */
int mm_gemm_resid_ln(const void* a, int64_t a_ld, const void* w, int64_t w_ld, int32_t rows, int32_t k, int32_t n,
                     const float* bias, float* x, const float* gamma, const float* beta, float eps, void* h_op,
                     float* h_f32, int32_t dtype, void* stream);

/* The same with dropout on the sub-layer output before the residual add (training forward: fairseq
 * TransformerEncoderLayer's dropout_module after out_proj / fc2): x_out <- x + dropout(a W^T + bias); mask as
 * mm_dropout at element index row * n + column. */
int mm_gemm_resid_ln_drop(const void* a, int64_t a_ld, const void* w, int64_t w_ld, int32_t rows, int32_t k, int32_t n,
                          const float* bias, const float* x, float* x_out, const float* gamma, const float* beta,
                          float eps, void* h_op, float* h_f32, float drop_p, uint64_t seed, const uint64_t* seed_dev,
                          uint32_t site, int32_t dtype, void* stream);

/* Same with the updated residual stream written to x_out instead of over x (x_out == x: in place): the training-step
 * forward keeps every sub-layer's input for the backward pass. */
int mm_gemm_resid_ln_out(const void* a, int64_t a_ld, const void* w, int64_t w_ld, int32_t rows, int32_t k, int32_t n,
                         const float* bias, const float* x, float* x_out, const float* gamma, const float* beta,
                         float eps, void* h_op, float* h_f32, int32_t dtype, void* stream);

/* LayerNorm over the last dim (eps 1e-5, affine), fp32 in -> 16-bit operand out and/or fp32 out.
 * Replaces F.layer_norm in fairseq TransformerEncoderLayer / final encoder LayerNorm and
 * image_pre_norm_module (mm_s2s_transformer.py:595).  dim in {256, 512, 768, 1024}. */
int mm_layernorm(const float* x, const float* gamma, const float* beta, int64_t rows, int32_t dim, void* out_op,
                 float* out_f32, int32_t dtype, float eps, void* stream);

/* Decoder input embedding: out[b, t] = scale * table[tokens[b, t]] + pos_table[position], position as fairseq
 * utils.make_positions (padding_idx + running count of non-pad tokens; pad tokens take row padding_idx = zeros).
 * fairseq TransformerDecoderBase.extract_features_scriptable (embed_scale * embed_tokens + embed_positions), the
 * input side of the S2UT unit decoder called at mm_s2s_transformer.py:693-696.  fp32 in, fp32 out [batch*length, dim]. */
int mm_embed_tokens(const int64_t* tokens, int32_t padding_idx, const float* table, int32_t vocab, float scale,
                    const float* pos_table, int32_t pos_rows, int32_t batch, int32_t length, int32_t dim, float* out,
                    void* stream);

/* Label-smoothed cross entropy terms of the unit logits: per row nll = -log_softmax(logits)[target] and
 * smooth = -sum_v log_softmax(logits)[v] (0 for rows whose target is padding_idx), plus their sums in sums[0..1]
 * (deterministic order).  The caller forms fairseq's loss = (1 - eps - eps_i) * sum_nll + eps_i * sum_smooth with
 * eps_i = eps / (vocab - 1)  (fairseq label_smoothed_nll_loss, used by the reference's criterion,
 * criterions/speech_to_speech_criterion.py:58-72).  logits fp32 [rows, ld], target int64 [rows]. */
int mm_label_smoothed_nll(const float* logits, int64_t ld, int32_t vocab, const int64_t* target, int32_t padding_idx,
                          int64_t rows, float* row_nll, float* row_smooth, float* sums, void* stream);

/* LayerNorm of rows gathered from a device-resident 16-bit feature store (x_dtype = MM_DTYPE_F16 / MM_DTYPE_BF16):
 * output row r = LayerNorm(x[index[r / rows_per_index] * rows_per_index + r % rows_per_index]); index == NULL: rows in
 * order.  This is image_pre_norm_module (mm_s2s_transformer.py:595) applied to the batch that the reference's
 * ImageDataset.__getitem__ (data/speech_to_speech_dataset.py:56-65) would have collated from host memory: the
 * features stay on the GPU and are read once, in 16 bit.  Output is the 16-bit GEMM operand only. */
int mm_layernorm_gather(const void* x, int32_t x_dtype, const int64_t* index, int32_t rows_per_index,
                        const float* gamma, const float* beta, int64_t rows, int32_t dim, void* out_op, int32_t dtype,
                        float eps, void* stream);

/* Multi-head self-attention core (fairseq MultiheadAttention: softmax_fp32(q k^T + key-padding mask) v).
 * qkv: [B*T, qkv_ld] 16-bit, q (pre-scaled by head_dim^-0.5) at columns [0, d), k at [d, 2d), v at [2d, 3d)
 * (exactly what the QKV projection GEMM writes); seq_lens [B] int32 valid keys; out [B*T, out_ld] 16-bit.
 * head_dim must be 64.  T <= 256 runs the persistent warp-specialised kernel, longer sequences the chunked one. */
int mm_self_attention(const void* qkv, int64_t qkv_ld, const int32_t* seq_lens, int32_t batch, int32_t seq,
                      int32_t heads, void* out, int64_t out_ld, int32_t dtype, void* stream);

/* General form: queries, keys and values from three tensors (or column blocks of one), q_len query rows and kv_len
 * key rows per batch element, kv_lens[b] valid keys (NULL: all), optional causal mask (key j visible to query i iff
 * j <= i; needs q_len == kv_len).  q is expected pre-scaled by head_dim^-0.5.  Rows are [batch][len][ld] with the
 * heads' 64-wide column blocks starting at *_col0.  Serves the S2UT decoder's causal self-attention and its
 * encoder attention (fairseq TransformerDecoderLayerBase.forward: self_attn / encoder_attn), i.e. the first
 * consumer of this path's output (mm_s2s_transformer.py:693-696).  Any lengths; runs the chunked online-softmax
 * kernel. */
int mm_attention(const void* q, int64_t q_ld, int32_t q_col0, int32_t q_len, const void* k, int64_t k_ld, int32_t k_col0,
                 const void* v, int64_t v_ld, int32_t v_col0, int32_t kv_len, const int32_t* kv_lens, int32_t batch,
                 int32_t heads, int32_t causal, void* out, int64_t out_ld, int32_t dtype, void* stream);

/* scores[r, k] = -inf where key_mask[r / rows_per_seq][k] != 0, in place (the reference's key-padding masked_fill,
 * fuse.py:88-91), for the training forward: the backward pass recomputes the probabilities from the kept scores, so
 * the mask has to live in them.  scores fp32 [rows, ld]; key_mask uint8 [rows / rows_per_seq][mask_ld]. */
int mm_mask_scores(float* scores, int64_t ld, int64_t rows, int32_t n_keys, const uint8_t* key_mask, int64_t mask_ld,
                   int32_t rows_per_seq, void* stream);

/* Fused (flash-style) speech -> image attention: out = softmax(q k^T + key mask) v for ONE head of width d_model
 * (SelectiveAttention.forward, mm_s2ut/models/fuse.py:80-113, built with num_heads = 1 at
 * mm_s2s_transformer.py:132-137; MultimodalAttention, fuse.py:145-167, with the learned bias_k / bias_v stored as the
 * last key / value row).  Scores and probabilities stay in TMEM: nothing but q, k, v is read and only out (and the
 * optional per-row log-sum-exp) is written.  q [batch * q_len, q_ld] 16-bit, pre-scaled by d_model^-0.5; k / v
 * [batch][kv_len][ld] 16-bit with the d_model-wide blocks at *_col0 (they may be the two halves of one K|V projection
 * output; kv_batch_stride elements between utterances, 0 = kv_len * ld); key_mask optional [batch][mask_ld] uint8,
 * non-zero = key masked out (the reference's masked_fill(-inf), fuse.py:88-91); out [batch * q_len, out_ld] 16-bit;
 * lse optional [batch * q_len] fp32.  d_model must be a multiple of 256. */
int mm_cross_attention(const void* q, int64_t q_ld, int32_t q_len, const void* k, int64_t k_ld, int32_t k_col0,
                       const void* v, int64_t v_ld, int32_t v_col0, int32_t kv_len, int64_t kv_batch_stride,
                       const uint8_t* key_mask, int64_t mask_ld, int32_t batch, int32_t d_model, void* out,
                       int64_t out_ld, float* lse, int32_t dtype, void* stream);

/* The same two kernels with one more output for the training step: lse [batch][heads][queries] fp32 = natural-log
 * sum-exp of every query row's (masked) scores, from which the backward pass rebuilds the probabilities without the
 * scores ever having been stored (mm_attention_bwd_scores).  lse == NULL: identical to the calls above. */
int mm_self_attention_lse(const void* qkv, int64_t qkv_ld, const int32_t* seq_lens, int32_t batch, int32_t seq,
                          int32_t heads, void* out, int64_t out_ld, float* lse, int32_t dtype, void* stream);
int mm_attention_lse(const void* q, int64_t q_ld, int32_t q_col0, int32_t q_len, const void* k, int64_t k_ld,
                     int32_t k_col0, const void* v, int64_t v_ld, int32_t v_col0, int32_t kv_len, const int32_t* kv_lens,
                     int32_t batch, int32_t heads, int32_t causal, void* out, int64_t out_ld, float* lse, int32_t dtype,
                     void* stream);

/* Attention backward, score side (what PyTorch autograd does for fairseq's MultiheadAttention in the reference's
 * training step): P = exp(q k^T - lse) with the key-length / causal masks, dS = P o (dO v^T - rowsum(dO o O)), written
 * as 16-bit [batch * heads][..][..] matrices with row stride pd_ld and matrix stride pd_bs (elements), for the
 * dV = P^T dO, dK = dS^T q, dQ = dS k contractions that follow (mm_gemm).  Arguments as mm_attention; dout / out are
 * the gradient and the forward result of the attention output [batch * q_len, heads * 64]; lse as written by
 * mm_self_attention_lse / mm_attention_lse.  Scores, dP and the softmax backward stay in TMEM / registers. */
int mm_attention_bwd_scores(const void* q, int64_t q_ld, int32_t q_col0, int32_t q_len, const void* k, int64_t k_ld,
                            int32_t k_col0, const void* v, int64_t v_ld, int32_t v_col0, int32_t kv_len,
                            const int32_t* kv_lens, int32_t batch, int32_t heads, int32_t causal, const void* dout,
                            int64_t do_ld, const void* out, int64_t o_ld, const float* lse, void* probs, void* dscores,
                            int64_t pd_ld, int64_t pd_bs, int32_t dtype, void* stream);

/* Row softmax for the speech->image attention (fuse.py:88-111): scores fp32 [rows, ld_in] -> probabilities
 * 16-bit [rows, ld_out]; columns [n_keys, ld_out) written as 0.  key_mask: optional [n_seqs, n_keys] uint8
 * (1 = padded key -> -inf); rows_per_seq maps a row to its sequence. */
int mm_softmax_rows(const float* scores, int64_t ld_in, int64_t rows, int32_t n_keys, const uint8_t* key_mask,
                    int32_t rows_per_seq, void* probs, int64_t ld_out, int32_t dtype, void* stream);

/* fp32 -> 16-bit operand conversion (weights / image features); n elements. */
int mm_convert_f32(const float* x, void* out, int64_t n, int32_t dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Training-step variant (BASELINE configs[2]): backward of the same path + optimizer.  The reference gets all of this
 * from PyTorch autograd over the modules named above (fairseq_cli/train -> task.train_step -> loss.backward();
 * optimizer = fairseq Adam, scripts/textless/1_train.sh) -- these are the `_bwd` twins SURVEY.md 8(b) asks for.
 * Every dense contraction of the backward pass (dgrad = dY W, wgrad = dY^T X) runs on mm_gemm; the entry points below
 * are the memory-bound pieces between them.
 *
 * mm_pack_t: 16-bit operand copies of `in` ([batches][rows][cols], fp32 or 16-bit, batch z at
 *   (z / nb1) * in_bs0 + (z % nb1) * in_bs1), multiplied by `scale`, optionally zeroed where mask <= 0 (ReLU backward):
 *   out_n [..][rows][n_ld] straight and/or out_t [..][cols][t_ld] TRANSPOSED, whose columns [rows, t_cols_pad) are
 *   written as 0 (split-K padding of the wgrad GEMMs).  Also splits / merges attention heads (nb1 = heads).
 * mm_rowsum: out[r] (+)= sum_c in[r, c]  (bias gradient from the transposed output gradient).
 * mm_colsum: partials [mm_colsum_blocks(rows)][cols] = per-512-row-chunk column sums of a 16-bit [rows, cols] matrix
 *   (bias gradients; rows with r % period >= valid are skipped when period > 0); summed by mm_reduce_partials.
 * mm_reduce_partials: out[i] (+)= sum_s part[s * stride + i]  (split-K partials, LayerNorm parameter partials).
 * mm_layernorm_bwd: dx = resid + LN'(dy) (dx / resid optional; dx_op: 16-bit copy of dx), partials [mm_layernorm_bwd_blocks()][2][dim] =
 *   per-block (sum dy * xhat, sum dy).
 * mm_softmax_bwd: P = softmax(scores[:, :valid]), dscores = P o (dprobs - rowsum(P o dprobs)), both 16-bit; dprobs is
 *   fp32 or (dprobs_is_op) 16-bit with its own leading dimension,
 *   rows [batch][rows_per_batch] of which the first valid_rows are processed (0: all), valid = kv_lens[batch / heads]
 *   (NULL: n_keys), further limited to the query's own index + 1 when causal; columns [valid, ld_out) = 0.
 * mm_glu_bwd: pre fp32 [rows, 2n] = (a | b), dy fp32 [rows, n] -> dpre 16-bit [rows, 2n]   (F.glu backward, x scale).
 * mm_gate_bwd: selective gate backward (mm_s2s_transformer.py:612-618): z = pre-sigmoid gate incl. bias ->
 *   dz 16-bit [B*T, d], dcat fp32 [B*T, 2d] = (dres g | dres (1-g)); dres is T x B x C.
 * mm_tbc_to_btc: T x B x C fp32 -> token-major fp32.
 * mm_col2im_k5s2: input gradient of Conv1d(k=5, stride 2, pad 2) from the per-window gradient [B, T_out, 5*C].
 * mm_grad_clip_coef: norm_coef[0] = ||grad_scale * grad||_2, norm_coef[1] = grad_scale * min(1, max_norm / (norm + 1e-6))
 *   (fairseq clip_grad_norm_; max_norm <= 0: no clipping); partials: mm_sumsq_blocks() floats.  norm_coef has 8 floats:
 *   [0] norm, [1] multiplier (outputs); with dev_hyper != 0 the kernels read the per-step hyper-parameters from it
 *   instead of their by-value arguments, so that a captured CUDA graph can be replayed with new values:
 *   [2] Adam step_size = lr sqrt(1-b2^t)/(1-b1^t), [3] weight_decay * lr, [4] grad_scale, [5] max_norm.
 *   (mm_adam takes them from the device when step == 0.)  extra_norm (optional): extra_norm[0] = the already scaled
 *   norm of another gradient buffer (e.g. the decoder's), combined into a joint norm before the clip coefficient.
 * mm_adam: fairseq.optim.adam.Adam.step on a flat fp32 buffer; the gradient is multiplied by norm_coef[1] (NULL: 1);
 *   param_op (optional): 16-bit copy of the updated parameters, written in the same pass (the GEMM operand copies).
 * --------------------------------------------------------------------------------------------- */
int mm_pack_t(const void* in, int32_t in_is_f32, int64_t in_ld, int64_t in_bs0, int64_t in_bs1, int32_t nb1,
              const void* mask, int64_t mask_ld, int32_t rows, int32_t cols, int32_t batches, float scale, void* out_n,
              int64_t n_ld, int64_t n_bs0, int64_t n_bs1, void* out_t, int64_t t_ld, int64_t t_bs0, int64_t t_bs1,
              int32_t t_cols_pad, int32_t dtype, void* stream);
int mm_rowsum(const void* in, int64_t ld, int32_t rows, int32_t cols, float* out, int32_t accumulate, int32_t dtype,
              void* stream);
int mm_reduce_partials(const float* part, int32_t n_partials, int64_t stride, int64_t n, float* out, int32_t accumulate,
                       void* stream);
/* Several mm_reduce_partials in one launch (the backward pass defers a layer's reductions and runs them together). */
/* Attention backward in one kernel for self-attention over sequences of up to 256 positions (the encoder at 10 s
 * utterances): dq | dk | dv from q | k | v, dO, O and lse with S, dP, P and dS never leaving the SM (TMEM / shared
 * memory).  qkv [batch * seq_len, qkv_ld] and dqkv [batch * seq_len, dqkv_ld] hold the heads' 64-wide column blocks of
 * q, k, v (and their gradients) at q_col0 / k_col0 / v_col0; q pre-scaled by head_dim^-0.5 and dq scaled likewise, as
 * in mm_attention_bwd_scores + mm_heads_gemm, which this replaces when seq_len <= 256.  kv_lens: valid keys per
 * sequence or NULL; dout / out: gradient and forward result of the attention output [batch * seq_len, heads * 64]; lse
 * [batch][heads][seq_len] from mm_self_attention_lse. */
int mm_attention_bwd_fused(const void* qkv, int64_t qkv_ld, int32_t q_col0, int32_t k_col0, int32_t v_col0,
                           int32_t seq_len, const int32_t* kv_lens, int32_t batch, int32_t heads, const void* dout,
                           int64_t do_ld, const void* out, int64_t o_ld, const float* lse, void* dqkv, int64_t dqkv_ld,
                           int32_t dtype, void* stream);
/* The two with attention dropout (fairseq MultiheadAttention dropout_module on the probabilities, --attention-dropout
 * 0.1 in scripts/textless/1_train.sh:112) generated INSIDE the kernels: the forward multiplies P by keep / (1 - p)
 * before the P V product (the softmax denominator and lse stay un-dropped), the backward regenerates the same mask
 * (counter-based, as mm_dropout, element index ((b heads + h) Tp + q) Tp + k with Tp = seq rounded up to 64).
 * mm_self_attention_drop: any length (the single-chunk kernel for 129 .. 256 positions, else the chunked one);
 * mm_attention_bwd_fused_drop: up to 256 positions (beyond: mm_attention_bwd_general_drop). */
int mm_self_attention_drop(const void* qkv, int64_t qkv_ld, const int32_t* seq_lens, int32_t batch, int32_t seq,
                           int32_t heads, void* out, int64_t out_ld, float* lse, float drop_p, uint64_t seed,
                           const uint64_t* seed_dev, uint32_t site, int32_t dtype, void* stream);
int mm_attention_bwd_fused_drop(const void* qkv, int64_t qkv_ld, int32_t q_col0, int32_t k_col0, int32_t v_col0,
                                int32_t seq_len, const int32_t* kv_lens, int32_t batch, int32_t heads, const void* dout,
                                int64_t do_ld, const void* out, int64_t o_ld, const float* lse, void* dqkv,
                                int64_t dqkv_ld, float drop_p, uint64_t seed, const uint64_t* seed_dev, uint32_t site,
                                int32_t dtype, void* stream);

/* The same for any query / key length, a causal mask and q / k|v (and dq / dk|dv) in different tensors: the unit
 * decoder's causal self-attention and its encoder attention (fairseq TransformerDecoderLayer under autograd), the
 * encoder beyond 256 positions.  Query tiles are processed in pairs; the dk / dv partial sums of the earlier pairs wait
 * in `scratch` (fp32, mm_attention_bwd_general_scratch_floats(kv_len) elements, caller-owned).  Arguments as
 * mm_attention_bwd_scores; dq [batch * q_len, dq_ld], dk / dv [batch * kv_len, ld] with the heads' column blocks at
 * *_col0. */
int64_t mm_attention_bwd_general_scratch_floats(int32_t kv_len);
int mm_attention_bwd_general(const void* q, int64_t q_ld, int32_t q_col0, int32_t q_len, const void* k, int64_t k_ld,
                             int32_t k_col0, const void* v, int64_t v_ld, int32_t v_col0, int32_t kv_len,
                             const int32_t* kv_lens, int32_t batch, int32_t heads, int32_t causal, const void* dout,
                             int64_t do_ld, const void* out, int64_t o_ld, const float* lse, void* dq, int64_t dq_ld,
                             int32_t dq_col0, void* dk, int64_t dk_ld, int32_t dk_col0, void* dv, int64_t dv_ld,
                             int32_t dv_col0, float* scratch, int32_t dtype, void* stream);
/* Attention dropout inside the general kernels (the unit decoder's two attentions under --attention-dropout,
 * scripts/textless/1_train.sh:112; the encoder beyond 256 positions): mm_attention_drop is mm_attention_lse with
 * P o keep / (1 - p) entering the P V product, mm_attention_bwd_general_drop regenerates the mask.  Element index of
 * (b, h, query q, key k) in the counter-based mask: ((b heads + h) Lp + q) Tp + k, Lp / Tp = q_len / kv_len rounded up
 * to 64 (the layout mm_softmax_dropout_bwd uses on stored [batch * heads][Lp][Tp] scores). */
int mm_attention_drop(const void* q, int64_t q_ld, int32_t q_col0, int32_t q_len, const void* k, int64_t k_ld,
                      int32_t k_col0, const void* v, int64_t v_ld, int32_t v_col0, int32_t kv_len, const int32_t* kv_lens,
                      int32_t batch, int32_t heads, int32_t causal, void* out, int64_t out_ld, float* lse, float drop_p,
                      uint64_t seed, const uint64_t* seed_dev, uint32_t site, int32_t dtype, void* stream);
int mm_attention_bwd_general_drop(const void* q, int64_t q_ld, int32_t q_col0, int32_t q_len, const void* k, int64_t k_ld,
                                  int32_t k_col0, const void* v, int64_t v_ld, int32_t v_col0, int32_t kv_len,
                                  const int32_t* kv_lens, int32_t batch, int32_t heads, int32_t causal, const void* dout,
                                  int64_t do_ld, const void* out, int64_t o_ld, const float* lse, void* dq, int64_t dq_ld,
                                  int32_t dq_col0, void* dk, int64_t dk_ld, int32_t dk_col0, void* dv, int64_t dv_ld,
                                  int32_t dv_col0, float* scratch, float drop_p, uint64_t seed, const uint64_t* seed_dev,
                                  uint32_t site, int32_t dtype, void* stream);

/* Attention backward, output side (autograd of fairseq's MultiheadAttention: dV = P^T dO, dK = dS^T q, dQ = dS k): per
 * sequence b and head h (head_dim 64)
 *     out[b][r][out_col0 + 64 h + c] = scale * sum_j A_bh[r, j] * w[b][j][w_col0 + 64 h + c]      r < rows, j < k
 * a holds the [rows, k] matrices of all (b, h) as mm_attention_bwd_scores wrote them: a_transposed == 0: element
 * (r, j) at a[bh * a_bs + r * a_ld + j]; != 0: at a[bh * a_bs + j * a_ld + r] (P^T, dS^T without a transposed copy).
 * w [batch][k tokens][w_ld] and out [batch][rows][out_ld] are token-major tensors with the heads' 64-wide column blocks
 * starting at *_col0 (q | k | v and their gradient), *_bs elements between sequences.  16-bit in and out. */
int mm_heads_gemm(const void* a, int64_t a_ld, int64_t a_bs, int32_t a_transposed, const void* w, int64_t w_ld,
                  int64_t w_bs, int32_t w_col0, void* out, int64_t out_ld, int64_t out_bs, int32_t out_col0,
                  int32_t rows, int32_t k, int32_t batch, int32_t heads, float scale, int32_t dtype, void* stream);

/* Gradient exchange over NVLink peer memory (the reference's data-parallel training, scripts/textless/1_train.sh:105-125,
 * exchanges gradients through fairseq's DDP all-reduce).  One process per GPU:
 *   mm_ipc_get_handle    CUDA IPC handle of the allocation holding `ptr` + ptr's byte offset inside it
 *   mm_ipc_open_handle   map another process's allocation (peer access enabled lazily); mm_ipc_close_handle unmaps it
 *   mm_p2p_allreduce_f32 bufs[p] = rank p's n-element fp32 buffer as mapped here (bufs[rank] = the local one): rank r
 *                        sums slice r of all buffers (rank order) and stores the result into slice r of all buffers.
 *                        The caller brackets the launch with barriers over the ranks (mm_p2p_barrier); bufs may point
 *                        at any common offset inside the mapped buffers (a gradient bucket). */
#define MM_P2P_MAX_RANKS 8
int mm_ipc_get_handle(const void* ptr, uint8_t* handle64, int64_t* offset);
int mm_ipc_open_handle(const uint8_t* handle64, void** mapped_base);
int mm_ipc_close_handle(void* mapped_base);
int mm_p2p_allreduce_f32(float* const* bufs, int32_t world, int32_t rank, int64_t n, void* stream);
/* Barrier over the ranks as a kernel (graph-capturable, no shared memory): flags[p] = rank p's array of
 * MM_P2P_MAX_RANKS uint32 (zero-initialised once) as mapped here, epoch = a local uint32 counter.  Work enqueued before
 * the barrier on any rank is visible to work enqueued after it on every rank.  Every rank must issue the same sequence
 * of barriers. */
int mm_p2p_barrier(unsigned int* const* flags, unsigned int* epoch, int32_t world, int32_t rank, void* stream);
/* The same exchange on 16-bit gradients (the reference recipe trains with --fp16, scripts/textless/1_train.sh:125, so
 * fairseq's all-reduce moves 16-bit gradients): mm_p2p_pack_bf16 rounds the fp32 gradients into a bf16 staging buffer
 * of n_pad >= n elements (a multiple of 8; the tail is zeroed), mm_p2p_allreduce_bf16 runs the two-shot exchange on the
 * ranks' staging buffers (fp32 accumulation in rank order, one rounding of the sum; n a multiple of 8), and
 * mm_p2p_unpack_bf16 widens the result -- identical on every rank -- back into the fp32 buffer.  Half the NVLink bytes. */
int mm_p2p_pack_bf16(const float* src, void* stage, int64_t n, int64_t n_pad, void* stream);
int mm_p2p_allreduce_bf16(void* const* bufs, int32_t world, int32_t rank, int64_t n, void* stream);
int mm_p2p_unpack_bf16(const void* stage, float* dst, int64_t n, void* stream);

/* Grouped weight gradient (autograd of nn.Linear inside fairseq's TransformerEncoderLayer / MultiheadAttention under
 * `loss.backward()`, scripts/textless/1_train.sh): for every group g
 *     out_g[n, k] (+)= sum_t dy_g[t, n] * x_g[t, k]        t = 0 .. tokens-1, fp32 accumulation, fp32 output
 * in ONE launch: the 256 x 256 output tiles of all groups share the persistent grid, so no token split (and no
 * partial sums) is needed to fill the GPU.  dy_g [tokens, n_out] and x_g [tokens, k_in] are the 16-bit tensors the
 * backward pass already holds (row strides dy_ld / x_ld, elements); out_g is the fp32 gradient [n_out, k_in] with row
 * stride out_ld.  accumulate != 0 adds to out (gradient accumulation over micro-batches). */
#define MM_WGRAD_MAX_GROUPS 64
typedef struct mm_wgrad_group {
  const void* dy;
  const void* x;
  float* out;
  int64_t dy_ld, x_ld, out_ld;
  int32_t n_out, k_in;
  float* bias;          /* optional: bias[n] (+)= sum_t dy[t, n]; k_in == 0 (x, out NULL) computes only this */
  int64_t tokens;       /* this group's token count; 0: the call's `tokens`.  Groups of different lengths (the encoder
                         * layers' B*T tokens, the image-side K / V projections' B*577) share one launch: the host
                         * orders the tiles by cost so that the CTA pairs finish together */
} mm_wgrad_group;
int mm_wgrad_grouped(const mm_wgrad_group* groups, int32_t count, int64_t tokens, int32_t accumulate, int32_t dtype,
                     void* stream);

/* Input gradient of a Linear that follows a LayerNorm, the LayerNorm backward and the residual add in ONE kernel (autograd
 * of `x = residual + sublayer(layer_norm(x))` in fairseq's TransformerEncoderLayer / TransformerDecoderLayer, d_model 512):
 *     dh = dy w                      dy [rows, k] 16-bit, w [k, 512] = the Linear's weight [out, in] as stored
 *     g += LayerNorm'(dh; x, gamma)  x [rows, 512] fp32 = the LayerNorm's input, g [rows, 512] fp32 in place
 *     g_op = 16-bit(g) (x keep / (1 - p) of dropout site `site` when drop_p > 0: the gradient entering the next branch)
 *     partials [mm_gemm_ln_bwd_partial_rows(rows)][2][512] = (dgamma, dbeta) partial sums (sum the rows with
 *     mm_reduce_partials_many).  dh stays in TMEM; replaces mm_gemm(MM_EPI_F32) + mm_layernorm_bwd_drop. */
int mm_gemm_ln_bwd_partial_rows(int64_t rows);
int mm_gemm_ln_bwd(const void* dy, int64_t dy_ld, const void* w, int64_t w_ld, int64_t rows, int32_t k, const float* x,
                   const float* gamma, float eps, float* g, void* g_op, float* partials, float drop_p, uint64_t seed,
                   const uint64_t* seed_dev, uint32_t site, int32_t dtype, void* stream);

typedef struct mm_reduce_job {
  const float* part;
  float* out;
  int64_t stride, n;
  int32_t n_partials, accumulate;
} mm_reduce_job;
int mm_reduce_partials_many(const mm_reduce_job* jobs, int32_t count, void* stream);
int mm_layernorm_bwd_blocks(void);
int mm_layernorm_bwd(const float* x, const float* gamma, const float* dy, int64_t rows, int32_t dim, float eps,
                     const float* resid, float* dx, float* partials, void* dx_op, int32_t dtype, void* stream);
/* The same with the dropout mask of the sub-layer branch the gradient enters next burnt into its 16-bit copy:
 * dx_op = (dx o keep) / (1 - drop_p) (mask as mm_dropout, element index row * dim + column); dx itself (the residual
 * path) stays unmasked. */
int mm_layernorm_bwd_drop(const float* x, const float* gamma, const float* dy, int64_t rows, int32_t dim, float eps,
                          const float* resid, float* dx, float* partials, void* dx_op, float drop_p, uint64_t seed,
                          const uint64_t* seed_dev, uint32_t site, int32_t dtype, void* stream);
int mm_softmax_bwd(const float* scores, const void* dprobs, int32_t dprobs_is_op, int64_t ld_dprobs, int64_t ld_in,
                   int64_t rows, int32_t rows_per_batch,
                   int32_t n_keys, const int32_t* kv_lens, int32_t heads, void* probs, void* dscores, int64_t ld_out,
                   int32_t valid_rows, int32_t causal, int32_t dtype, void* stream);
/* The same with attention dropout on the probabilities (fairseq MultiheadAttention dropout_module, fuse.py:111): the mask
 * m = keep(seed + *seed_dev, site, row * ld_out + k) / (1 - p) is regenerated, probs = P m (the matrix that multiplied V),
 * dscores = P o (dprobs m - rowsum(P o dprobs m)).  dprobs == NULL and dscores == NULL: the training FORWARD (softmax +
 * dropout -> probs). */
int mm_softmax_dropout_bwd(const float* scores, const void* dprobs, int32_t dprobs_is_op, int64_t ld_dprobs, int64_t ld_in,
                           int64_t rows, int32_t rows_per_batch, int32_t n_keys, const int32_t* kv_lens, int32_t heads,
                           void* probs, void* dscores, int64_t ld_out, int32_t valid_rows, int32_t causal, float drop_p,
                           uint64_t seed, const uint64_t* seed_dev, uint32_t site, int32_t dtype, void* stream);
/* Backward of mm_label_smoothed_nll summed over rows (fairseq label_smoothed_nll_loss, reduce=True), times grad_scale:
 * dlogits 16-bit [rows, ld_out] (columns >= vocab and padding rows are 0) -- the A operand of the tied output
 * projection's dgrad / wgrad.  mm_embed_tokens_bwd: table_grad[v] += scale * sum of dx[row] over rows whose token is v
 * (one block per vocabulary row, rows summed in order: deterministic; the padding row is skipped). */
int mm_label_smoothed_nll_bwd(const float* logits, int64_t ld, int32_t vocab, const int64_t* target, int32_t padding_idx,
                              int64_t rows, float epsilon, float grad_scale, void* dlogits, int64_t ld_out, int32_t dtype,
                              void* stream);
int mm_embed_tokens_bwd(const int64_t* tokens, int32_t padding_idx, const float* dx, int64_t rows, int32_t dim,
                        float scale, float* table_grad, int32_t vocab, void* stream);
int mm_colsum_blocks(int32_t rows);
int mm_colsum(const void* in, int64_t ld, int32_t rows, int32_t cols, int32_t period, int32_t valid, float* partials,
              int32_t dtype, void* stream);
int mm_glu_bwd(const float* pre, const float* dy, int64_t rows, int32_t n, float scale, void* dpre, int32_t dtype,
               void* stream);
int mm_gate_bwd(const float* z, const float* dres_tbc, const float* text, const float* attn, int32_t batch, int32_t seq,
                int32_t dim, void* dz, float* dcat, int32_t dtype, void* stream);
int mm_tbc_to_btc(const float* in_tbc, int32_t batch, int32_t seq, int32_t dim, float* out, void* stream);
int mm_col2im_k5s2(const float* dcol, int32_t batch, int32_t t_out, int32_t t_in, int32_t channels, float* dx,
                   void* stream);
/* Element-wise dropout of the training forward / backward (fairseq FairseqDropout at the encoder's embedding, residual
 * and activation sites; SA_image_dropout, mm_s2s_transformer.py:596): out = (resid ? resid : 0) + (keep ? x / (1-p) : 0),
 * keep = f(seed + *seed_dev, site, element index) -- a pure function, regenerated in the backward pass (seed_dev:
 * optional device-resident per-step seed, so that CUDA-graph replays draw fresh masks).  x / out fp32 or 16-bit. */
int mm_dropout(const void* x, int32_t x_is_f32, const float* resid, void* out, int64_t n, float p, uint64_t seed,
               const uint64_t* seed_dev, uint32_t site, int32_t dtype, void* stream);
int mm_sumsq_blocks(void);
int mm_grad_clip_coef(const float* grad, int64_t n, float grad_scale, float max_norm, float* partials, float* norm_coef,
                      int32_t dev_hyper, const float* extra_norm, void* stream);
int mm_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
            float beta2, float eps, float weight_decay, int32_t step, const float* norm_coef, void* param_op,
            int32_t dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMS2UT_B200_H_ */
