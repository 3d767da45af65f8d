/* unidisc_b200 — C ABI of the B200-native UniDisc hot path (libunidisc_b200.so).
 *
 * The reference (alexanderswerdlow/unidisc @ 01b6125c) is pure Python and has no FFI: its operator boundary is the
 * Python class `models.dit.DIT` plus the pure-tensor `Diffusion` methods.  This header is the C-level
 * seam those Python entry points bind through `ctypes` (unidisc_b200/_lib.py); each function names the reference
 * code it replaces (file:line relative to the reference repo).  Conventions:
 *   - every pointer is a DEVICE pointer unless stated; plain sizes; no torch types; `stream` is a cudaStream_t;
 *   - return value 0 = launched OK, non-zero = CUDA/argument error (message on stderr); nothing is synchronised;
 *   - bf16 = raw __nv_bfloat16 bits, tokens / indices are int64 like the reference's tensors;
 *   - row-major everywhere, `ld*` = leading dimension in ELEMENTS.
 */
#ifndef UNIDISC_B200_H
#define UNIDISC_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- library info ------------------------------------------------------------------------------------------ */
int ud_abi_version(void);          /* bumped on any signature change */
int ud_device_sm_count(void);

/* ---- GEMM family (tcgen05 / TMEM / TMA) -------------------------------------------------------------------- */
enum {
    UD_EPI_BF16 = 0,       /* C bf16 = acc (+ bias)                                     nn.Linear, dit.py:642,887,919,1091 */
    UD_EPI_BF16_GELU = 1,  /* C bf16 = u = acc + bias ; aux bf16 = gelu_tanh(u)         mlp.0 + nn.GELU("tanh"), dit.py:917-919 */
    UD_EPI_BF16_DGELU = 2, /* C bf16 = acc * gelu_tanh'(aux)  (aux = saved u)           autograd of the above */
    UD_EPI_F32 = 3,        /* C fp32 = acc ; aux (optional) = fp32 device scalar += sum C^2   weight gradient (+ its share of the
                              gradient norm of clip_grad_norm_, model.py:1518; needs M > 128) */
    UD_EPI_F32_ACC = 4,    /* C fp32 += acc                                             weight gradient accumulation (.grad +=) */
    UD_EPI_BF16_SCALED = 5 /* C bf16 = bf16(bf16(acc) * alpha), alpha = *aux (device fp32)   weight gradient written straight in the wire
                              format of torch's bf16 compress hook (`buffer.to(bf16).div_(world)`, main.py:643-648): the DDP
                              bucket is all-reduced in place, no fp32 store + compression pass.  (ta, tb) = (1, 1) only */
};
/* C[M,N] = sum_k A(m,k) B(n,k).  ta=0: A is [M,K] (lda);  ta=1: A is [K,M] (lda).  tb=0: B is [N,K];  tb=1: B is [K,N].
 * Supported (ta,tb): (0,0) forward, (0,1) dgrad, (1,1) wgrad.  A,B bf16, 16-byte aligned, lda/ldb multiples of 8.
 * bias: bf16 [N] or NULL.  bn_hint: 128 / 256 / 0 (auto).  One kernel family: persistent CTA pairs (cta_group::2), any M. */
int ud_gemm_bf16(int ta, int tb, int M, int N, int K, const void* A, long long lda, const void* B, long long ldb, void* C,
                 long long ldc, int epi, const void* bias, void* aux, long long ld_aux, int bn_hint, void* stream);

/* ---- time conditioning (config.time_conditioning; dit.py:966-967 adaLN chunks, 258-268/301-304 modulate_fused, 229-253
 * bias_dropout_add_scale with modality).  Passed as NULL by every shipped training config.  The adaLN Linear's bf16 output
 * [B, ld] is consumed in place: shift/scale modulate the h output of the fused norm kernels on rows with sel = 1
 * (image tokens; every token when the batch holds no image token), gate multiplies the dropped branch on img rows (text rows get
 * the plain branch).  Backward accumulates the per-sample gradients with atomics into fp32 [B, ld_d] buffers. */
typedef struct ud_adaln {
    const uint8_t* sel;        /* [rows] */
    const uint8_t* img;        /* [rows] */
    const void* shift;         /* bf16, element (b, j) at [b * ld + j]; NULL = no modulation of h */
    const void* scale;
    const void* gate;          /* bf16; NULL = branch is not gated (pre_residual_norm under sandwich normalisation) */
    long long ld;
    int tokens_per_sample;     /* sample of row r = r / tokens_per_sample */
    float* d_shift;            /* backward only */
    float* d_scale;
    float* d_gate;
    long long ld_d;
} ud_adaln;

/* ---- embedding + first RMSNorm ------------------------------------------------------------------------------
 * x = E[ids] + Emod[modality] (+ Ecount[ordinal] where ordinal >= 0)  (dit.py:1375,1406,163-167);
 * h = bf16(rms(x) * w)  (dit.py:95-100,971).  rows = B*N.  ordinal (int32 [rows], from ud_interleaved_prep) / Ecount
 * (`img_count_embedding`, fp32 [16,D]) are NULL outside interleaved batches. */
int ud_embed_rmsnorm_fwd(const int64_t* ids, const int64_t* modality, const float* E, const float* Emod, const float* w,
                         float* x, void* h_bf16, float* rstd, int rows, int D, float eps, const int* ordinal,
                         const float* Ecount, const ud_adaln* tc /* NULL = no time conditioning */, void* stream);
/* dE[ids] += g ; dEmod[modality] += g ; dEcount[ordinal] += g   (g = gradient wrt x, fp32 [rows,D]) */
int ud_embed_bwd(const int64_t* ids, const int64_t* modality, const float* g, float* dE, float* dEmod, int rows, int D,
                 long long hot_id /* id accumulated per CTA (the mask token), -1 = none */, const int* ordinal,
                 float* dEcount, void* stream);

/* ---- interleaved-batch preparation (data.require_sample_ids; dit.py:122-191,1421-1443, tensor_utils.py:4-44) ----------
 * One launch replaces the reference's Python loops over image / sample blocks.  modality, sample_ids: int64 [B,N].
 * cos_tab/sin_tab: fp32 [rows_total, hd2] = the 1-D text table at row txt_off and the 2-D tables of the 256/1024/2304/4096
 * -token image blocks at off256..off4096 (dit.py:1208-1212).  Outputs per token: cos/sin [B*N, hd2] and the image ordinal
 * (index into img_count_embedding, -1 = none).  Image block = maximal run of modality != 0; blocks whose size has no
 * table keep cos = sin = 0 and ordinal -1 (as in the reference); text tokens use (pos - start of their sample_id run);
 * pad runs (sample_id < 0) stay 0.  scratch: int32 [B, 4N]. */
int ud_interleaved_prep(const int64_t* modality, const int64_t* sample_ids, int B, int N, const float* cos_tab,
                        const float* sin_tab, int hd2, int txt_off, int off256, int off1024, int off2304, int off4096,
                        float* cos_out, float* sin_out, int* ordinal_out, int* scratch, void* stream);

/* ---- fused "sandwich" norm + residual + next pre-norm -------------------------------------------------------
 * x_out = x_in + dropout_p(bf16(rms(a)) * w_a) ;  h = bf16(rms(x_out) * w_n)
 * = pre_residual_norm/post_ff_norm + residual (dit.py:993-994,1024-1031) fused with the following norm2 / next block's
 * norm1 / norm_final (dit.py:971,1025,1089).  a: bf16 [rows,D] branch output; x_in/x_out fp32; w_* fp32 [D].
 * p_drop > 0: training-mode dropout of the branch (bias_dropout_add_scale, dit.py:229-253; model.dropout) with an in-kernel
 * Philox4x32-10 mask keyed by (seed, offset, row, column group); backward regenerates the mask from the same triple. */
int ud_norm_residual_fwd(const void* a_bf16, const float* x_in, const float* w_a, const float* w_n, float* x_out,
                         void* h_bf16, float* rstd_a, float* rstd_x, int rows, int D, float eps, float p_drop, uint64_t seed,
                         uint64_t offset, const ud_adaln* tc, void* stream);
/* the keep-scales (0 or 1/(1-p), fp32 [rows,D]) the two kernels above/below use for (p_drop, seed, offset) — test hook */
int ud_dropout_scales(float* out, int rows, int D, float p_drop, uint64_t seed, uint64_t offset, void* stream);
/* backward of the above.  g_out: fp32 grad wrt x_out from the residual stream (may be NULL = 0); dh: bf16 grad wrt h.
 * Writes g_in (fp32 total grad wrt x_out == grad wrt x_in), da (bf16 grad wrt a), and atomically accumulates
 * dw_n += sum_rows dh * xhat,  dw_a += sum_rows g * bf16(rms(a)),  and (if db_a != NULL) db_a += sum_rows da — the bias
 * gradient of the Linear that produced `a` (mlp.2.bias, dit.py:919), saving a separate column-sum pass. */
int ud_norm_residual_bwd(const float* g_out, const void* dh_bf16, const float* x_out, const float* rstd_x, const float* w_n,
                         const void* a_bf16, const float* rstd_a, const float* w_a, float* g_in, void* da_bf16,
                         float* dw_n, float* dw_a, float* db_a, int rows, int D, float p_drop, uint64_t seed, uint64_t offset,
                         const ud_adaln* tc, void* stream);
/* backward of the first norm only: g_in = g_out + rms_bwd(dh) ; dw += ... */
int ud_rmsnorm_bwd(const float* g_out, const void* dh_bf16, const float* x, const float* rstd, const float* w, float* g_in,
                   float* dw, int rows, int D, const ud_adaln* tc, void* stream);

/* ---- q/k LayerNorm (over the full hidden dim) + RoPE ---------------------------------------------------------
 * qk_out[:, 0:D] = rope(bf16(LN(q))), qk_out[:, D:2D] = rope(bf16(LN(k)))   (dit.py:680-682, 724-726,
 * standalone_rotary.py:14-31).  qkv: bf16 [rows,3D] (three h d); cos/sin: fp32 [rows, hd/2] per-token tables
 * (dit.py:1419-1458); stats: fp32 [rows,4] = (mean_q, rstd_q, mean_k, rstd_k) saved for backward. */
int ud_qk_ln_rope_fwd(const void* qkv_bf16, const float* gq, const float* bq, const float* gk, const float* bk,
                      const float* cos, const float* sin, void* qk_out_bf16, float* stats, int rows, int D, int head_dim,
                      float eps, void* stream);
/* backward: dqk bf16 [rows,2D] (grads wrt rotated q,k) -> dqkv[:, 0:2D] (bf16, ld 3D); accumulates dgq,dbq,dgk,dbk. */
int ud_qk_ln_rope_bwd(const void* dqk_bf16, const void* qkv_bf16, const float* stats, const float* gq, const float* gk,
                      const float* cos, const float* sin, void* dqkv_bf16, float* dgq, float* dbq, float* dgk, float* dbk,
                      int rows, int D, int head_dim, void* stream);

/* ---- attention (tcgen05, bidirectional softmax(QK^T/sqrt(hd))V, optional document mask) ----------------------
 * replaces torch SDPA / FlexAttention (dit.py:775-829).  q,k: bf16 [B*N, ldqk] (head h at column h*hd), v: bf16 with ldv,
 * o: bf16 [B*N, ldo]; lse: fp32 [B,H,N] (natural log, scaled scores).  sample_ids: int64 [B,N] or NULL
 * (mask = same id and id != -1, model_utils.py:740-771).  With sample_ids every CTA visits only the key tiles whose id range
 * overlaps its query tile's (the block skipping of the reference's BlockMask): cost ~ sum(len_i^2), not N^2. */
int ud_attn_fwd(const void* q, const void* k, long long ldqk, const void* v, long long ldv, void* o, long long ldo, float* lse,
                const int64_t* sample_ids, int B, int N, int H, int head_dim, float scale, void* stream);
/* partial-query attention against a K/V cache (inference: dit.py:588-614 `update_kv_cache`, 793-812 image-K/V cache of the
 * FlexAttention path): Nq query tokens per sample attend to Nk key/value tokens, no mask.  Every operand is a token-major bf16
 * matrix with row pitch ld* and a per-sample stride *_bs (both in elements; *_bs = 0 means rows * ld, i.e. dense), so sub-ranges
 * of a sequence (the text rows of q, the image rows of a cache) are addressed in place.  lse: fp32 [B,H,Nq]. */
int ud_attn_fwd_kv(const void* q, long long ldq, long long q_bs, const void* k, long long ldk, long long k_bs, const void* v,
                   long long ldv, long long v_bs, void* o, long long ldo, long long o_bs, float* lse, int B, int Nq, int Nk, int H,
                   int head_dim, float scale, void* stream);
/* backward: writes dq,dk (bf16, ld lddqk) and dv (bf16, ld lddv).  delta: fp32 scratch of 2*B*H*N floats
 * ([0] rowsum(dO*O), [1] lse in the exp2 domain; both produced by the kernel's own first pass). */
int ud_attn_bwd(const void* q, const void* k, long long ldqk, const void* v, long long ldv, const void* o, const void* d_o,
                long long ldo, const float* lse, float* delta, void* dq, void* dk, long long lddqk, void* dv, long long lddv,
                const int64_t* sample_ids, int B, int N, int H, int head_dim, float scale, void* stream);

/* ---- column sums (bias gradients): db[n] += sum_m dY[m,n] ---------------------------------------------------- */
int ud_colsum_bf16(const void* dY, long long ld, float* db, int M, int N, void* stream);

/* ---- SUBS parameterisation + NLL (model.py:621-658, 967) ------------------------------------------------------
 * Per token row r: logits get -1e6 at mask_index and at wrong-modality vocabulary; log-softmax (fp32 maths);
 * carry-over rule for unmasked xt.  logp[r] = log p(x0[r]); lse[r] saved for backward.  logits: bf16 [rows, ldv]. */
int ud_subs_nll_fwd(const void* logits_bf16, long long ldv, const int64_t* xt, const int64_t* x0, const int64_t* modality,
                    float* logp, float* lse, int rows, int V, int text_vocab, int mask_index, void* stream);
/* dlogits (bf16, in place over logits, all ldv columns written) = dlogp[r] * (onehot(x0) - softmax) on masked rows, else 0 */
int ud_subs_nll_bwd(void* logits_bf16, long long ldv, const int64_t* xt, const int64_t* x0, const int64_t* modality,
                    const float* lse, const float* dlogp, int rows, int V, int text_vocab, int mask_index, void* stream);
/* full SUBS log-probs, for API parity with `_subs_parameterization` (model.py:621-658): out fp32 or bf16 [rows, V] */
int ud_subs_logprobs(const void* logits_bf16, long long ldv, const int64_t* xt /*may be NULL*/, const int64_t* modality,
                     void* out, int out_is_bf16, long long ldo, int rows, int V, int text_vocab, int mask_index, void* stream);

/* ---- absorbing-state masking q_xt (model.py:439,579) --------------------------------------------------------
 * xt = (rand < move_chance[b]) ? mask_index : x.  rand: fp32 [B,N] supplied (parity mode) or NULL -> Philox4x32-10
 * with (seed, offset).  move (uint8 [B,N], optional) receives the move mask. */
int ud_q_xt(const int64_t* x, const float* move_chance, const float* rand, uint64_t seed, uint64_t offset, int64_t mask_index,
            int64_t* xt, uint8_t* move, int B, int N, void* stream);

/* ---- categorical / absorbing samplers (model_utils.py:95-97, model_eval.py:2042-2104) ------------------------
 * out[r] = argmax_v p[r,v] / (1e-10 - log(u[r,v] + 1e-10)).  probs fp32 [R,V]; u fp32 [R,V] or NULL (Philox). */
int ud_sample_categorical(const float* probs, long long ldp, const float* u, uint64_t seed, uint64_t offset, int64_t* out,
                          int R, int V, void* stream);
/* x' = copy_flag*x + (1-copy_flag)*sample(q), q = p_x0*(mc_t-mc_s), q[mask]=mc_s.  p_x0 fp32 [B*N,V]. */
int ud_ddpm_update_probs(const int64_t* x, const float* p_x0, long long ldp, const float* u, uint64_t seed, uint64_t offset,
                         const float* mc_t, const float* mc_s, int64_t mask_index, int64_t* out, int B, int N, int V,
                         void* stream);
/* fused fast path: raw bf16 logits -> SUBS softmax -> absorbing update, never materialising p_x0.  cond (optional):
 * second logits tensor + per-sample cfg weight w[b] for (1+w)*c - w*u (model_eval.py:1812). */
int ud_ddpm_update_logits(const int64_t* x, const void* logits_bf16, const void* logits_uncond_bf16, long long ldv,
                          const float* cfg_w, const int64_t* modality, const float* u, uint64_t seed, uint64_t offset,
                          const float* mc_t, const float* mc_s, int64_t mask_index, int text_vocab, int64_t* out, int B, int N,
                          int V, void* stream);

/* MaskGIT step (model_eval.py:3045-3114) in two launches: (1) per masked token row, one vocabulary pass: SUBS softmax p,
 * pred = torch.multinomial(p, 1) (= argmax_v p_v / E_v with E ~ Exp(1): ATen's single-draw multinomial), conf = log p_pred +
 * r_temp * gumbel * t (float64, as the reference's np.random.gumbel draw promotes it; -inf on unmasked rows);  (2) per sample,
 * the num_unmask[b] (clamped to the number of masked tokens) most confident rows take their prediction, everything else keeps
 * x.  e_noise fp32 [B*N, V] + gumbel fp64 [B*N] supplied = bit-parity mode; both NULL = in-kernel Philox draws.
 * t: fp32 [B]; num_unmask: int32 [B] (= schedule[:, step]); pred (int64 [B*N]) and conf (fp64 [B*N]) are outputs/scratch. */
int ud_maskgit_update(const int64_t* x, const void* logits_bf16, const void* logits_uncond_bf16, long long ldv,
                      const float* cfg_w, const int64_t* modality, const float* e_noise, const double* gumbel, uint64_t seed,
                      uint64_t offset, const float* t, float r_temp, const int* num_unmask, int64_t mask_index, int text_vocab,
                      int64_t* pred, double* conf, int64_t* out, int B, int N, int V, void* stream);
/* arg-max over the vocabulary of the SUBS log-probs with carry-over (x where xt != mask): the noise-removal pass of `_sample`
 * (model_eval.py:2440-2446) without materialising [rows, V] log-probs */
int ud_subs_argmax(const void* logits_bf16, long long ldv, const int64_t* xt, const int64_t* modality, int64_t* out, int rows,
                   int V, int text_vocab, int mask_index, void* stream);

/* ---- optimizer / DDP helpers ----------------------------------------------------------------------------------
 * fused AdamW (torch.optim.AdamW semantics, model_setup.py:385-424) over a flat fp32 buffer, also emitting the bf16
 * shadow copy the GEMMs read.  grad_scale multiplies the gradient (clip coefficient); step is 1-based. */
int ud_adamw_step(float* p, const float* g, float* m, float* v, void* p_bf16, long long n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int step, const float* grad_scale /*device scalar or NULL*/,
                  int max_ctas /* > 0 caps the grid (side-stream use next to GEMMs) */, void* stream);
int ud_cast_f32_to_bf16(const float* src, void* dst_bf16, long long n, void* stream);
/* sum of squares of a flat fp32 buffer, accumulated into out[0] (out must be zeroed by the caller) */
int ud_sumsq_f32(const float* g, long long n, float* out, int max_ctas, void* stream);
/* DDP bf16 compress hook (torch default_hooks._compress_hook): dst = bf16(bf16(g) / world) ; and decompress.
 * max_ctas > 0 caps the grid so the side-stream copies leave the SMs to the backward GEMMs they overlap with. */
int ud_grad_pack_bf16(const float* g, void* dst_bf16, long long n, float inv_world, int max_ctas, void* stream);
int ud_grad_unpack_bf16(const void* src_bf16, float* g, long long n, int max_ctas, void* stream);
/* decompress and, on the way, sumsq[0] += sum of squares of the fp32 gradients written (the all-reduced gradient's share of the
 * norm clip_grad_norm_ needs, model.py:1518) */
int ud_grad_unpack_bf16_sumsq(const void* src_bf16, float* g, long long n, int max_ctas, float* sumsq, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UNIDISC_B200_H */
