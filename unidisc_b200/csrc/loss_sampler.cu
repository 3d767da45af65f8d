// HBM-bound vocabulary-row kernels (sm_100a): SUBS parameterisation + NLL (forward/backward), absorbing-state masking
// q_xt, Gumbel-arg-max categorical sampler and the fused absorbing (ddpm / ddpm_cache) update.
//
// One CTA per token row, 16-byte coalesced loads over the row's VALID vocabulary range only (text rows never touch the
// image vocabulary and vice versa: the reference adds -1e6 there, model.py:627-635, which contributes exactly 0 to the
// fp32 log-sum-exp), warp-shuffle + smem block reductions, one pass over HBM per row.
#include "common.cuh"
#include "unidisc_b200.h"

namespace ud {

static constexpr float NEG_INF_SUBS = -1000000.0f;  // self.neg_infinity (model.py:626)
static constexpr int ROW_THREADS = 256;

struct MaxSum { float m, s; };
UD_DEVINL void ms_add(MaxSum& a, float x) {
    if (x > a.m) { a.s = a.s * __expf(a.m - x) + 1.0f; a.m = x; }
    else a.s += __expf(x - a.m);
}
UD_DEVINL MaxSum ms_merge(MaxSum a, MaxSum b) {
    if (b.m > a.m) { MaxSum t = a; a = b; b = t; }
    if (b.m == -INFINITY) return a;
    a.s += b.s * __expf(b.m - a.m);
    return a;
}
UD_DEVINL MaxSum block_ms(MaxSum v, float* sm /*[2*32]*/) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        MaxSum t;
        t.m = __shfl_xor_sync(0xffffffffu, v.m, o);
        t.s = __shfl_xor_sync(0xffffffffu, v.s, o);
        v = ms_merge(v, t);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();  // protect sm from the previous use
    if (lane == 0) { sm[warp] = v.m; sm[32 + warp] = v.s; }
    __syncthreads();
    MaxSum r = {sm[0], sm[32]};
    for (int w = 1; w < nw; ++w) r = ms_merge(r, MaxSum{sm[w], sm[32 + w]});
    return r;
}

// valid vocabulary range of a row (model.py:627-635)
UD_DEVINL void valid_range(long long modality, int V, int text_vocab, int& lo, int& hi) {
    if (text_vocab <= 0) { lo = 0; hi = V; }
    else if (modality == 0) { lo = 0; hi = text_vocab; }
    else { lo = text_vocab; hi = V; }
}

// log-sum-exp of a bf16 row over [lo,hi) excluding column `skip`
UD_DEVINL MaxSum row_lse_bf16(const __nv_bfloat16* row, int lo, int hi, int skip) {
    MaxSum acc = {-INFINITY, 0.f};
    const int lo_al = min(hi, (lo + 7) & ~7), hi_al = max(lo_al, hi & ~7);
    for (int v = lo + threadIdx.x; v < lo_al; v += blockDim.x)
        if (v != skip) ms_add(acc, __bfloat162float(row[v]));
    for (int v = lo_al + threadIdx.x * 8; v < hi_al; v += blockDim.x * 8) {
        const uint4 t = ldg_stream(row + v);
        const float f[8] = {bf16lo(t.x), bf16hi(t.x), bf16lo(t.y), bf16hi(t.y), bf16lo(t.z), bf16hi(t.z), bf16lo(t.w), bf16hi(t.w)};
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (v + i != skip) ms_add(acc, f[i]);
    }
    for (int v = hi_al + threadIdx.x; v < hi; v += blockDim.x)
        if (v != skip) ms_add(acc, __bfloat162float(row[v]));
    return acc;
}

// ------------------------------------------------------------------------------------------------
// SUBS NLL forward: logp[r] = log p_theta(x0[r] | xt)   (model.py:621-658 + gather at model.py:967)
// ------------------------------------------------------------------------------------------------
__global__ void subs_nll_fwd_kernel(const __nv_bfloat16* __restrict__ logits, long long ldv, const int64_t* __restrict__ xt,
                                    const int64_t* __restrict__ x0, const int64_t* __restrict__ modality,
                                    float* __restrict__ logp, float* __restrict__ lse_out, int rows, int V, int text_vocab,
                                    int mask_index) {
    __shared__ float sm[64];
    for (int r = blockIdx.x; r < rows; r += gridDim.x) {
        const long long xtr = xt[r], x0r = x0[r];
        if (xtr != mask_index) {  // carry-over: the row is one-hot at xt (model.py:646-656)
            if (threadIdx.x == 0) { logp[r] = (x0r == xtr) ? 0.0f : NEG_INF_SUBS; lse_out[r] = 0.0f; }
            continue;
        }
        int lo, hi;
        valid_range(modality[r], V, text_vocab, lo, hi);
        const __nv_bfloat16* row = logits + (long long)r * ldv;
        const MaxSum t = block_ms(row_lse_bf16(row, lo, hi, mask_index), sm);
        if (threadIdx.x == 0) {
            const float lse = t.m + logf(t.s);
            float l = __bfloat162float(row[x0r]);
            if (x0r < lo || x0r >= hi || x0r == mask_index) l += NEG_INF_SUBS;
            logp[r] = l - lse;
            lse_out[r] = lse;
        }
    }
}

// dlogits = dlogp[r] * (onehot(x0) - softmax) on masked rows, 0 elsewhere; in place; every one of the ldv columns written
__global__ void subs_nll_bwd_kernel(__nv_bfloat16* __restrict__ logits, long long ldv, const int64_t* __restrict__ xt,
                                    const int64_t* __restrict__ x0, const int64_t* __restrict__ modality,
                                    const float* __restrict__ lse_in, const float* __restrict__ dlogp, int rows, int V,
                                    int text_vocab, int mask_index) {
    for (int r = blockIdx.x; r < rows; r += gridDim.x) {
        __nv_bfloat16* row = logits + (long long)r * ldv;
        const float g = dlogp[r];
        const bool active = (xt[r] == mask_index) && (g != 0.0f);
        int lo = 0, hi = 0;
        if (active) valid_range(modality[r], V, text_vocab, lo, hi);
        const float lse = lse_in[r];
        const int x0r = (int)x0[r];
        for (int v = threadIdx.x * 8; v < ldv; v += blockDim.x * 8) {
            uint4 o = make_uint4(0, 0, 0, 0);
            if (active && v + 8 > lo && v < hi) {
                const uint4 t = *reinterpret_cast<const uint4*>(row + v);
                const float f[8] = {bf16lo(t.x), bf16hi(t.x), bf16lo(t.y), bf16hi(t.y), bf16lo(t.z), bf16hi(t.z), bf16lo(t.w), bf16hi(t.w)};
                float d[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int c = v + i;
                    const bool valid = c >= lo && c < hi && c != mask_index;
                    d[i] = valid ? -g * __expf(f[i] - lse) : 0.0f;
                    if (c == x0r) d[i] += g;
                }
                o = make_uint4(pack_bf16x2(d[0], d[1]), pack_bf16x2(d[2], d[3]), pack_bf16x2(d[4], d[5]), pack_bf16x2(d[6], d[7]));
            } else if (active && x0r >= v && x0r < v + 8) {
                float d[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                d[x0r - v] = g;
                o = make_uint4(pack_bf16x2(d[0], d[1]), pack_bf16x2(d[2], d[3]), pack_bf16x2(d[4], d[5]), pack_bf16x2(d[6], d[7]));
            }
            *reinterpret_cast<uint4*>(row + v) = o;
        }
    }
}

// full SUBS log-prob rows (API parity with _subs_parameterization)
template <bool OUT_BF16>
__global__ void subs_logprobs_kernel(const __nv_bfloat16* __restrict__ logits, long long ldv, const int64_t* __restrict__ xt,
                                     const int64_t* __restrict__ modality, void* __restrict__ out, long long ldo, int rows,
                                     int V, int text_vocab, int mask_index) {
    __shared__ float sm[64];
    for (int r = blockIdx.x; r < rows; r += gridDim.x) {
        const __nv_bfloat16* row = logits + (long long)r * ldv;
        const bool carry = xt != nullptr && xt[r] != mask_index;
        const long long xtr = carry ? xt[r] : -1;
        int lo, hi;
        valid_range(modality[r], V, text_vocab, lo, hi);
        float lse = 0.f;
        if (!carry) {
            const MaxSum t = block_ms(row_lse_bf16(row, lo, hi, mask_index), sm);
            lse = t.m + logf(t.s);
        }
        for (int v = threadIdx.x; v < V; v += blockDim.x) {
            float o;
            if (carry) o = (v == xtr) ? 0.0f : NEG_INF_SUBS;
            else {
                float l = __bfloat162float(row[v]);
                if (v < lo || v >= hi || v == mask_index) l += NEG_INF_SUBS;
                o = l - lse;
            }
            if (OUT_BF16) reinterpret_cast<__nv_bfloat16*>(out)[(long long)r * ldo + v] = __float2bfloat16_rn(o);
            else reinterpret_cast<float*>(out)[(long long)r * ldo + v] = o;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// q_xt   (model.py:439,579)
// ------------------------------------------------------------------------------------------------
__global__ void q_xt_kernel(const int64_t* __restrict__ x, const float* __restrict__ move_chance, const float* __restrict__ rnd,
                            uint64_t seed, uint64_t offset, int64_t mask_index, int64_t* __restrict__ xt,
                            uint8_t* __restrict__ move, int B, int N) {
    const long long total = (long long)B * N;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(i / N);
        const float u = rnd ? rnd[i] : philox_uniform(seed, offset, (uint64_t)i);
        const bool mv = u < move_chance[b];
        xt[i] = mv ? mask_index : x[i];
        if (move) move[i] = mv ? 1 : 0;
    }
}

// ------------------------------------------------------------------------------------------------
// Gumbel arg-max:  argmax_v  q_v / (1e-10 - log(u_v + 1e-10))     (model_utils.py:95-97)
// ------------------------------------------------------------------------------------------------
struct ArgMax { float v; int i; };
UD_DEVINL void am_upd(ArgMax& a, float v, int i) {
    if (v > a.v || (v == a.v && i < a.i)) { a.v = v; a.i = i; }
}
UD_DEVINL ArgMax block_argmax(ArgMax a, float* smv, int* smi) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, a.v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, a.i, o);
        am_upd(a, ov, oi);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) { smv[warp] = a.v; smi[warp] = a.i; }
    __syncthreads();
    ArgMax r = {smv[0], smi[0]};
    for (int w = 1; w < nw; ++w) am_upd(r, smv[w], smi[w]);
    return r;
}
UD_DEVINL float gumbel_norm(float u) { return 1e-10f - logf(u + 1e-10f); }

// MODE 0: plain categorical over probs.  MODE 1: absorbing update (q = p*(mc_t-mc_s), q[mask] = mc_s, copy-through).
template <int MODE>
__global__ void sample_probs_kernel(const int64_t* __restrict__ x, const float* __restrict__ probs, long long ldp,
                                    const float* __restrict__ u, uint64_t seed, uint64_t offset, const float* __restrict__ mc_t,
                                    const float* __restrict__ mc_s, int64_t mask_index, int64_t* __restrict__ out, int R, int N,
                                    int V) {
    __shared__ float smv[32];
    __shared__ int smi[32];
    for (int r = blockIdx.x; r < R; r += gridDim.x) {
        float d = 1.0f, ms = 0.0f;
        if (MODE == 1) {
            const long long xr = x[r];
            if (xr != mask_index) {  // copy_flag (model_eval.py:2064,2093)
                if (threadIdx.x == 0) out[r] = xr;
                continue;
            }
            const int b = r / N;
            d = mc_t[b] - mc_s[b];
            ms = mc_s[b];
        }
        const float* pr = probs + (long long)r * ldp;
        const float* ur = u ? u + (long long)r * V : nullptr;
        ArgMax a = {-INFINITY, 0x7fffffff};
        for (int v = threadIdx.x; v < V; v += blockDim.x) {
            float q = pr[v];
            if (MODE == 1) q = (v == mask_index) ? ms : q * d;
            const float uu = ur ? ur[v] : philox_uniform(seed, offset, (uint64_t)r * V + v);
            am_upd(a, __fdiv_rn(q, gumbel_norm(uu)), v);
        }
        a = block_argmax(a, smv, smi);
        if (threadIdx.x == 0) out[r] = a.i;
    }
}

// Stage one row's valid vocabulary range [lo,hi) in shared memory as fp32 (CFG-combined when CFG) and return the row's
// log-sum-exp over it (mask column excluded): HBM is read exactly once per row.  Block-wide; `sm` = 64 floats of scratch.
template <bool CFG>
UD_DEVINL float stage_row_lse(float* srow, float* sm, const __nv_bfloat16* rc, const __nv_bfloat16* ru, float w, int lo, int hi,
                              int mask_index) {
    MaxSum acc = {-INFINITY, 0.f};
    __syncthreads();  // srow reuse across rows
    auto put = [&](int v, float lcv, float luv) {
        float l = lcv;
        if (CFG) l = __fsub_rn(__fmul_rn(1.0f + w, lcv), __fmul_rn(w, luv));     // model_eval.py:1812 (torch: two products, one subtraction)
        srow[v - lo] = l;
        if (v != mask_index) ms_add(acc, l);
    };
    const int lo_al = min(hi, (lo + 7) & ~7), hi_al = max(lo_al, hi & ~7);
    for (int v = lo + threadIdx.x; v < lo_al; v += blockDim.x)
        put(v, __bfloat162float(rc[v]), CFG ? __bfloat162float(ru[v]) : 0.f);
    for (int v = lo_al + threadIdx.x * 8; v < hi_al; v += blockDim.x * 8) {
        const uint4 a = ldg_stream(rc + v);
        uint4 bq = make_uint4(0, 0, 0, 0);
        if (CFG) bq = ldg_stream(ru + v);
        put(v + 0, bf16lo(a.x), bf16lo(bq.x)); put(v + 1, bf16hi(a.x), bf16hi(bq.x));
        put(v + 2, bf16lo(a.y), bf16lo(bq.y)); put(v + 3, bf16hi(a.y), bf16hi(bq.y));
        put(v + 4, bf16lo(a.z), bf16lo(bq.z)); put(v + 5, bf16hi(a.z), bf16hi(bq.z));
        put(v + 6, bf16lo(a.w), bf16lo(bq.w)); put(v + 7, bf16hi(a.w), bf16hi(bq.w));
    }
    for (int v = hi_al + threadIdx.x; v < hi; v += blockDim.x)
        put(v, __bfloat162float(rc[v]), CFG ? __bfloat162float(ru[v]) : 0.f);
    const MaxSum t = block_ms(acc, sm);      // (its barriers also publish srow to the whole block)
    return t.m + logf(t.s);
}

// Fused path with SUPPLIED noise (bit-parity mode): bf16 logits (+ optional CFG pair) -> SUBS softmax -> absorbing update
// in the reference's fp32 operation order: q_v = exp(l_v - lse) * (mc_t - mc_s), q_mask = mc_s, arg-max of
// q_v / (1e-10 - log(u_v + 1e-10))   (model_eval.py:2091-2094, model_utils.py:95-97).
template <bool CFG>
__global__ void ddpm_update_logits_kernel(const int64_t* __restrict__ x, const __nv_bfloat16* __restrict__ lc,
                                          const __nv_bfloat16* __restrict__ lu, long long ldv, const float* __restrict__ cfg_w,
                                          const int64_t* __restrict__ modality, const float* __restrict__ u, uint64_t seed,
                                          uint64_t offset, const float* __restrict__ mc_t, const float* __restrict__ mc_s,
                                          int64_t mask_index, int text_vocab, int64_t* __restrict__ out, int R, int N, int V) {
    extern __shared__ float srow[];
    __shared__ float sm[64];
    __shared__ float smv[32];
    __shared__ int smi[32];
    for (int r = blockIdx.x; r < R; r += gridDim.x) {
        const long long xr = x[r];
        if (xr != mask_index) {
            if (threadIdx.x == 0) out[r] = xr;
            continue;
        }
        const int b = r / N;
        int lo, hi;
        valid_range(modality[r], V, text_vocab, lo, hi);
        const float lse = stage_row_lse<CFG>(srow, sm, lc + (long long)r * ldv, CFG ? lu + (long long)r * ldv : nullptr,
                                             CFG ? cfg_w[b] : 0.f, lo, hi, (int)mask_index);
        const float d = __fsub_rn(mc_t[b], mc_s[b]), ms = mc_s[b];
        const float* ur = u ? u + (long long)r * V : nullptr;
        ArgMax a = {-INFINITY, 0x7fffffff};
        for (int v = lo + threadIdx.x; v < hi; v += blockDim.x) {
            if (v == mask_index) continue;
            const float uu = ur ? ur[v] : philox_uniform(seed, offset, (uint64_t)r * V + v);
            am_upd(a, __fdiv_rn(__fmul_rn(expf(__fsub_rn(srow[v - lo], lse)), d), gumbel_norm(uu)), v);
        }
        if (threadIdx.x == 0) {  // the mask column keeps probability mc_s (model_eval.py:2066,2092)
            const float uu = ur ? ur[mask_index] : philox_uniform(seed, offset, (uint64_t)r * V + mask_index);
            am_upd(a, __fdiv_rn(ms, gumbel_norm(uu)), (int)mask_index);
        }
        a = block_argmax(a, smv, smi);
        if (threadIdx.x == 0) out[r] = a.i;
    }
}

// ------------------------------------------------------------------------------------------------
// MaskGIT step (model_eval.py:3045-3114), supplied-noise (bit-parity) draw: per masked token row
//   pred = torch.multinomial(p, 1) = argmax_v p_v / E_v,  E ~ Exp(1)   (ATen multinomial's single-draw path: q.exponential_(1); argmax(p / q))
//   conf = log(p_pred) + r_temp * gumbel * t   (float64 like the reference, whose np.random.gumbel draw promotes the sum)
// with p = exp(SUBS log-softmax) computed in one vocabulary pass; conf = -inf on rows that are already unmasked.
// ------------------------------------------------------------------------------------------------
template <bool CFG>
__global__ void maskgit_draw_kernel(const int64_t* __restrict__ x, const __nv_bfloat16* __restrict__ lc,
                                    const __nv_bfloat16* __restrict__ lu, long long ldv, const float* __restrict__ cfg_w,
                                    const int64_t* __restrict__ modality, const float* __restrict__ e_noise,
                                    const double* __restrict__ gumbel, const float* __restrict__ t, float r_temp,
                                    int64_t mask_index, int text_vocab, int64_t* __restrict__ pred, double* __restrict__ conf,
                                    int R, int N, int V) {
    extern __shared__ float srow[];
    __shared__ float sm[64];
    __shared__ float smv[32];
    __shared__ int smi[32];
    for (int r = blockIdx.x; r < R; r += gridDim.x) {
        const long long xr = x[r];
        if (xr != mask_index) {
            if (threadIdx.x == 0) { pred[r] = xr; conf[r] = -INFINITY; }
            continue;
        }
        const int b = r / N;
        int lo, hi;
        valid_range(modality[r], V, text_vocab, lo, hi);
        const float lse = stage_row_lse<CFG>(srow, sm, lc + (long long)r * ldv, CFG ? lu + (long long)r * ldv : nullptr,
                                             CFG ? cfg_w[b] : 0.f, lo, hi, (int)mask_index);
        const float* er = e_noise + (long long)r * V;
        ArgMax a = {-INFINITY, 0x7fffffff};
        for (int v = lo + threadIdx.x; v < hi; v += blockDim.x) {
            if (v == mask_index) continue;
            am_upd(a, __fdiv_rn(expf(__fsub_rn(srow[v - lo], lse)), er[v]), v);
        }
        a = block_argmax(a, smv, smi);
        if (threadIdx.x == 0) {
            const float p = expf(__fsub_rn(srow[a.i - lo], lse));
            pred[r] = a.i;
            conf[r] = (double)logf(p) + ((double)r_temp * gumbel[r]) * (double)t[b];
        }
    }
}

// MaskGIT selection: per sample keep the num_unmask[b] most confident predictions (ties at the threshold are all kept, like
// the reference's `conf >= k-th largest`).  One CTA per sample; conf staged in shared memory; rank by counting.
__global__ void maskgit_select_kernel(const int64_t* __restrict__ x, const int64_t* __restrict__ pred,
                                      const double* __restrict__ conf, const int* __restrict__ num_unmask, int64_t mask_index,
                                      int64_t* __restrict__ out, int N) {
    extern __shared__ double sconf[];
    __shared__ int s_cnt;
    const int b = blockIdx.x;
    const int64_t* xb = x + (long long)b * N;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    int local = 0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        sconf[i] = conf[(long long)b * N + i];
        local += xb[i] == mask_index;
    }
    local = (int)warp_sum((float)local);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_cnt, local);
    __syncthreads();
    const int k = min(num_unmask[b], s_cnt);                        // model_eval.py:3069
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const double ci = sconf[i];
        int greater = 0;
        if (k > 0) {
            for (int j = 0; j < N; ++j) greater += sconf[j] > ci;
        }
        const bool sel = k > 0 && greater < k && ci == ci;          // conf >= k-th largest value  (NaN never selected)
        out[(long long)b * N + i] = sel ? pred[(long long)b * N + i] : xb[i];
    }
}

// arg-max of the SUBS log-probs with carry-over (noise-removal pass of `_sample`, model_eval.py:2440-2446): x where already
// unmasked, else argmax_v (l_v - lse) over the valid vocabulary (first index on ties, like torch.argmax)
__global__ void subs_argmax_kernel(const __nv_bfloat16* __restrict__ logits, long long ldv, const int64_t* __restrict__ xt,
                                   const int64_t* __restrict__ modality, int64_t* __restrict__ out, int rows, int V,
                                   int text_vocab, int mask_index) {
    __shared__ float sm[64];
    __shared__ float smv[32];
    __shared__ int smi[32];
    for (int r = blockIdx.x; r < rows; r += gridDim.x) {
        const long long xr = xt[r];
        if (xr != mask_index) {
            if (threadIdx.x == 0) out[r] = xr;
            continue;
        }
        int lo, hi;
        valid_range(modality[r], V, text_vocab, lo, hi);
        const __nv_bfloat16* row = logits + (long long)r * ldv;
        const MaxSum t = block_ms(row_lse_bf16(row, lo, hi, mask_index), sm);
        const float lse = t.m + logf(t.s);
        ArgMax a = {-INFINITY, 0x7fffffff};
        const int lo_al = min(hi, (lo + 7) & ~7), hi_al = max(lo_al, hi & ~7);
        for (int v = lo + threadIdx.x; v < lo_al; v += blockDim.x)
            if (v != mask_index) am_upd(a, __bfloat162float(row[v]) - lse, v);
        for (int v = lo_al + threadIdx.x * 8; v < hi_al; v += blockDim.x * 8) {      // second pass over the row: L2-resident
            const uint4 t = *reinterpret_cast<const uint4*>(row + v);
            const float f[8] = {bf16lo(t.x), bf16hi(t.x), bf16lo(t.y), bf16hi(t.y), bf16lo(t.z), bf16hi(t.z), bf16lo(t.w), bf16hi(t.w)};
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (v + i != mask_index) am_upd(a, f[i] - lse, v + i);
        }
        for (int v = hi_al + threadIdx.x; v < hi; v += blockDim.x)
            if (v != mask_index) am_upd(a, __bfloat162float(row[v]) - lse, v);
        a = block_argmax(a, smv, smi);
        if (threadIdx.x == 0) out[r] = a.i;
    }
}


// Philox ("fast") variant of the fused absorbing update: ONE pass over the row, ~10 instructions per logit, HBM-bound.
// The Gumbel arg-max of model_utils.py:95-97 is evaluated hierarchically and in the log domain (both exact in
// distribution): P(8-column group) ~ sum of its q_v, then P(v | group) ~ q_v, so the row pass needs ONE uniform per 8 logits:
//     group score = log2( sum_{v in group} exp(l_v) ) - log2(E),   E = -ln(u) ~ Exp(1)     (Gumbel-max over groups)
// computed in the same pass as the online log-sum-exp; the winning group is re-read once (16 bytes) and a token drawn
// inside it; the mask column (probability mc_s, model_eval.py:2066/2092) is compared at the end.
UD_DEVINL float u01_from_bits(uint32_t w) { return ((float)(w >> 8) + 0.5f) * (1.0f / 16777216.0f); }   // (0,1)
UD_DEVINL float lg2f(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
UD_DEVINL float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// MODE 0: absorbing update (out = next token).  MODE 1: MaskGIT draw (model_eval.py:3072-3076): out = pred ~ p (no "stay"
// option), conf = log p_pred + r_temp * Gumbel * t with an in-kernel Gumbel draw; mc_t then carries t[b], mc_s is unused.
template <bool CFG, int MODE>
__global__ void __launch_bounds__(256)
ddpm_update_logits_fast_kernel(const int64_t* __restrict__ x, const __nv_bfloat16* __restrict__ lc,
                               const __nv_bfloat16* __restrict__ lu, long long ldv, const float* __restrict__ cfg_w,
                               const int64_t* __restrict__ modality, uint64_t seed, uint64_t offset,
                               const float* __restrict__ mc_t, const float* __restrict__ mc_s, int64_t mask_index,
                               int text_vocab, int64_t* __restrict__ out, int R, int N, int V, float r_temp,
                               double* __restrict__ conf) {
    // ONE WARP PER ROW: no shared memory, no block barriers; 64 independent rows in flight per SM hide the HBM latency and
    // the short serial tail (final draw by lane 0).
    constexpr float L2E = 1.4426950408889634f, LN2_ = 0.6931471805599453f;
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(offset >> 32));
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    for (int r = gw; r < R; r += nw) {
        const long long xr = x[r];
        if (xr != mask_index) {
            if (lane == 0) { out[r] = xr; if (MODE == 1) conf[r] = -INFINITY; }
            continue;
        }
        const int b = r / N;
        int lo, hi;
        valid_range(modality[r], V, text_vocab, lo, hi);
        const __nv_bfloat16* rc = lc + (long long)r * ldv;
        const __nv_bfloat16* ru = CFG ? lu + (long long)r * ldv : nullptr;
        const float w = CFG ? cfg_w[b] : 0.f;
        float m2 = -INFINITY, ssum = 0.f;        // running max (log2 units) and sum of 2^(l2 - m2)
        float best = -INFINITY;                  // best group score (log2 units)
        int best_v = 0x7fffffff;
        const int g0 = lo & ~7;
        int it = 0;
        for (int base = g0; base < hi; base += 1024, ++it) {
            uint4 a[4], bq[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {        // 4 fully coalesced 16-byte loads in flight per lane
                const int v = base + k * 256 + lane * 8;
                a[k] = make_uint4(0, 0, 0, 0);
                bq[k] = make_uint4(0, 0, 0, 0);
                if (v < hi) {
                    a[k] = ldg_stream(rc + v);
                    if (CFG) bq[k] = ldg_stream(ru + v);
                }
            }
            const uint4 rn = philox4x32_10(make_uint4((uint32_t)r, (uint32_t)it, (uint32_t)lane, (uint32_t)offset), key);
            const uint32_t rnd[4] = {rn.x, rn.y, rn.z, rn.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int v = base + k * 256 + lane * 8;
                if (v >= hi) continue;
                float l2[8] = {bf16lo(a[k].x), bf16hi(a[k].x), bf16lo(a[k].y), bf16hi(a[k].y),
                               bf16lo(a[k].z), bf16hi(a[k].z), bf16lo(a[k].w), bf16hi(a[k].w)};
                if (CFG) {
                    const float f[8] = {bf16lo(bq[k].x), bf16hi(bq[k].x), bf16lo(bq[k].y), bf16hi(bq[k].y),
                                        bf16lo(bq[k].z), bf16hi(bq[k].z), bf16lo(bq[k].w), bf16hi(bq[k].w)};
#pragma unroll
                    for (int i = 0; i < 8; ++i) l2[i] = (1.0f + w) * l2[i] - w * f[i];     // model_eval.py:1812
                }
                float gm = -INFINITY;
                const bool edge = (v < lo) || (v + 8 > hi) || (mask_index >= v && mask_index < v + 8);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    l2[i] *= L2E;
                    if (edge) { const int c = v + i; if (c < lo || c >= hi || c == mask_index) l2[i] = -INFINITY; }
                    gm = fmaxf(gm, l2[i]);
                }
                if (gm == -INFINITY) continue;
                if (gm > m2) { ssum *= (m2 == -INFINITY) ? 0.f : ex2f(m2 - gm); m2 = gm; }
                float gsum = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) gsum += ex2f(l2[i] - m2);
                ssum += gsum;
                const float E = -LN2_ * lg2f(u01_from_bits(rnd[k]));
                const float sc = m2 + lg2f(gsum) - lg2f(E);
                if (sc > best) { best = sc; best_v = v; }
            }
        }
        // warp merge: log-sum-exp (log2 units) and arg-max over group scores
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float om = __shfl_xor_sync(0xffffffffu, m2, o), os = __shfl_xor_sync(0xffffffffu, ssum, o);
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int ov = __shfl_xor_sync(0xffffffffu, best_v, o);
            const float nm = fmaxf(m2, om);
            if (nm != -INFINITY) ssum = ssum * ex2f(m2 - nm) + os * ex2f(om - nm);
            m2 = nm;
            if (ob > best || (ob == best && ov < best_v)) { best = ob; best_v = ov; }
        }
        if (lane == 0) {
            const float lse2 = m2 + lg2f(ssum);                                   // log2 of the row's sum of exp(l)
            const uint4 fin = philox4x32_10(make_uint4((uint32_t)r, 0xffffffffu, 0u, (uint32_t)offset), key);
            float gscore = 0.f, stay = -INFINITY, chosen_l2 = -INFINITY;
            if (MODE == 0) {
                const float d = mc_t[b] - mc_s[b], msk = mc_s[b];
                gscore = best + lg2f(fmaxf(d, 0.f)) - lse2;                           // log2(q_group) + Gumbel (log2 units)
                stay = lg2f(fmaxf(msk, 0.f)) - lg2f(-LN2_ * lg2f(u01_from_bits(fin.x)));
            }
            int64_t res = mask_index;
            if (best_v != 0x7fffffff && !(stay > gscore)) {
                // draw the token inside the winning group: Gumbel-max over its (up to 8) valid columns
                const uint4 fin2 = philox4x32_10(make_uint4((uint32_t)r, 0xfffffffeu, 0u, (uint32_t)offset), key);
                const uint4 fin3 = philox4x32_10(make_uint4((uint32_t)r, 0xfffffffdu, 0u, (uint32_t)offset), key);
                const int v = best_v;
                const uint4 a4 = *reinterpret_cast<const uint4*>(rc + v);
                float l2[8] = {bf16lo(a4.x), bf16hi(a4.x), bf16lo(a4.y), bf16hi(a4.y), bf16lo(a4.z), bf16hi(a4.z), bf16lo(a4.w), bf16hi(a4.w)};
                if (CFG) {
                    const uint4 b4 = *reinterpret_cast<const uint4*>(ru + v);
                    const float f[8] = {bf16lo(b4.x), bf16hi(b4.x), bf16lo(b4.y), bf16hi(b4.y), bf16lo(b4.z), bf16hi(b4.z), bf16lo(b4.w), bf16hi(b4.w)};
#pragma unroll
                    for (int i = 0; i < 8; ++i) l2[i] = (1.0f + w) * l2[i] - w * f[i];
                }
                const uint32_t rr[8] = {fin2.x, fin2.y, fin2.z, fin2.w, fin3.x, fin3.y, fin3.z, fin3.w};
                float bs = -INFINITY;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int c = v + i;
                    if (c < lo || c >= hi || c == mask_index) continue;
                    const float sc = l2[i] * L2E - lg2f(-LN2_ * lg2f(u01_from_bits(rr[i])));
                    if (sc > bs) { bs = sc; res = c; chosen_l2 = l2[i] * L2E; }
                }
            }
            out[r] = res;
            if (MODE == 1) {
                const float gum = -logf(-logf(u01_from_bits(fin.x)));                 // standard Gumbel (np.random.gumbel's law)
                conf[r] = (double)((chosen_l2 - lse2) * LN2_) + ((double)r_temp * (double)gum) * (double)mc_t[b];
            }
        }
    }
}

static int rows_grid(int rows) {
    long long g = (long long)sm_count() * 8;
    return (int)(rows < g ? rows : g);
}

}  // namespace ud

using namespace ud;
#define STREAM(s) reinterpret_cast<cudaStream_t>(s)
#define CBF(p) reinterpret_cast<const __nv_bfloat16*>(p)
#define BF(p) reinterpret_cast<__nv_bfloat16*>(p)

extern "C" int ud_subs_nll_fwd(const void* logits, long long ldv, const int64_t* xt, const int64_t* x0, const int64_t* modality,
                               float* logp, float* lse, int rows, int V, int text_vocab, int mask_index, void* stream) {
    if (rows <= 0) return 0;
    if (ldv % 8 != 0) { fprintf(stderr, "unidisc_b200: subs_nll needs ldv %% 8 == 0\n"); return -1; }
    subs_nll_fwd_kernel<<<rows_grid(rows), ROW_THREADS, 0, STREAM(stream)>>>(CBF(logits), ldv, xt, x0, modality, logp, lse, rows, V, text_vocab, mask_index);
    UD_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int ud_subs_nll_bwd(void* logits, long long ldv, const int64_t* xt, const int64_t* x0, const int64_t* modality,
                               const float* lse, const float* dlogp, int rows, int V, int text_vocab, int mask_index,
                               void* stream) {
    if (rows <= 0) return 0;
    if (ldv % 8 != 0) { fprintf(stderr, "unidisc_b200: subs_nll needs ldv %% 8 == 0\n"); return -1; }
    subs_nll_bwd_kernel<<<rows_grid(rows), ROW_THREADS, 0, STREAM(stream)>>>(BF(logits), ldv, xt, x0, modality, lse, dlogp, rows, V, text_vocab, mask_index);
    UD_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int ud_subs_logprobs(const void* logits, long long ldv, const int64_t* xt, const int64_t* modality, void* out,
                                int out_is_bf16, long long ldo, int rows, int V, int text_vocab, int mask_index, void* stream) {
    if (rows <= 0) return 0;
    if (ldv % 8 != 0) return -1;
    if (out_is_bf16)
        subs_logprobs_kernel<true><<<rows_grid(rows), ROW_THREADS, 0, STREAM(stream)>>>(CBF(logits), ldv, xt, modality, out, ldo, rows, V, text_vocab, mask_index);
    else
        subs_logprobs_kernel<false><<<rows_grid(rows), ROW_THREADS, 0, STREAM(stream)>>>(CBF(logits), ldv, xt, modality, out, ldo, rows, V, text_vocab, mask_index);
    UD_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int ud_q_xt(const int64_t* x, const float* move_chance, const float* rand, uint64_t seed, uint64_t offset,
                       int64_t mask_index, int64_t* xt, uint8_t* move, int B, int N, void* stream) {
    const long long total = (long long)B * N;
    if (total <= 0) return 0;
    long long blocks = (total + 255) / 256;
    if (blocks > (long long)sm_count() * 8) blocks = (long long)sm_count() * 8;
    q_xt_kernel<<<(int)blocks, 256, 0, STREAM(stream)>>>(x, move_chance, rand, seed, offset, mask_index, xt, move, B, N);
    UD_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int ud_sample_categorical(const float* probs, long long ldp, const float* u, uint64_t seed, uint64_t offset,
                                     int64_t* out, int R, int V, void* stream) {
    if (R <= 0) return 0;
    sample_probs_kernel<0><<<rows_grid(R), ROW_THREADS, 0, STREAM(stream)>>>(nullptr, probs, ldp, u, seed, offset, nullptr, nullptr, -1, out, R, 1, V);
    UD_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int ud_ddpm_update_probs(const int64_t* x, const float* p_x0, long long ldp, const float* u, uint64_t seed,
                                    uint64_t offset, const float* mc_t, const float* mc_s, int64_t mask_index, int64_t* out,
                                    int B, int N, int V, void* stream) {
    const int R = B * N;
    if (R <= 0) return 0;
    sample_probs_kernel<1><<<rows_grid(R), ROW_THREADS, 0, STREAM(stream)>>>(x, p_x0, ldp, u, seed, offset, mc_t, mc_s, mask_index, out, R, N, V);
    UD_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int ud_ddpm_update_logits(const int64_t* x, const void* logits, const void* logits_uncond, long long ldv,
                                     const float* cfg_w, const int64_t* modality, const float* u, uint64_t seed,
                                     uint64_t offset, const float* mc_t, const float* mc_s, int64_t mask_index, int text_vocab,
                                     int64_t* out, int B, int N, int V, void* stream) {
    const int R = B * N;
    if (R <= 0) return 0;
    int widest = V;
    if (text_vocab > 0) widest = text_vocab > V - text_vocab ? text_vocab : V - text_vocab;
    const size_t smem = (size_t)widest * sizeof(float);
    if (smem > 200 * 1024) { fprintf(stderr, "unidisc_b200: vocabulary range too wide for the fused sampler (%d)\n", widest); return -1; }
    const bool cfg = logits_uncond != nullptr;
    if (u == nullptr) {   // Philox noise: single-pass log-domain kernel, no shared-memory staging
        if (ldv % 8 != 0) { fprintf(stderr, "unidisc_b200: fused sampler needs ldv %% 8 == 0\n"); return -1; }
        long long want = ((long long)R * 32 + 255) / 256;          // one warp per row, 8 warps per CTA
        const int grid = (int)(want < (long long)sm_count() * 8 ? want : (long long)sm_count() * 8);
        if (cfg)
            ddpm_update_logits_fast_kernel<true, 0><<<grid, 256, 0, STREAM(stream)>>>(x, CBF(logits), CBF(logits_uncond), ldv, cfg_w, modality, seed, offset, mc_t, mc_s, mask_index, text_vocab, out, R, N, V, 0.f, nullptr);
        else
            ddpm_update_logits_fast_kernel<false, 0><<<grid, 256, 0, STREAM(stream)>>>(x, CBF(logits), nullptr, ldv, nullptr, modality, seed, offset, mc_t, mc_s, mask_index, text_vocab, out, R, N, V, 0.f, nullptr);
        UD_CUDA_CHECK(cudaGetLastError());
        return 0;
    }
    static bool attr[2] = {false, false};
    if (cfg) {
        if (!attr[1]) { UD_CUDA_CHECK(cudaFuncSetAttribute(ddpm_update_logits_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr[1] = true; }
        ddpm_update_logits_kernel<true><<<rows_grid(R), 512, smem, STREAM(stream)>>>(x, CBF(logits), CBF(logits_uncond), ldv, cfg_w, modality, u, seed, offset, mc_t, mc_s, mask_index, text_vocab, out, R, N, V);
    } else {
        if (!attr[0]) { UD_CUDA_CHECK(cudaFuncSetAttribute(ddpm_update_logits_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr[0] = true; }
        ddpm_update_logits_kernel<false><<<rows_grid(R), 512, smem, STREAM(stream)>>>(x, CBF(logits), nullptr, ldv, nullptr, modality, u, seed, offset, mc_t, mc_s, mask_index, text_vocab, out, R, N, V);
    }
    UD_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int ud_maskgit_update(const int64_t* x, const void* logits, const void* logits_uncond, long long ldv, const float* cfg_w,
                                 const int64_t* modality, const float* e_noise, const double* gumbel, uint64_t seed,
                                 uint64_t offset, const float* t, float r_temp, const int* num_unmask, int64_t mask_index,
                                 int text_vocab, int64_t* pred, double* conf, int64_t* out, int B, int N, int V, void* stream) {
    const int R = B * N;
    if (R <= 0) return 0;
    if (ldv % 8 != 0) { fprintf(stderr, "unidisc_b200: maskgit needs ldv %% 8 == 0\n"); return -1; }
    if ((e_noise == nullptr) != (gumbel == nullptr)) { fprintf(stderr, "unidisc_b200: maskgit needs both noise tensors or neither\n"); return -1; }
    const bool cfg = logits_uncond != nullptr;
    if (e_noise == nullptr) {        // in-kernel Philox: single-pass warp-per-row draw
        long long want = ((long long)R * 32 + 255) / 256;
        const int grid = (int)(want < (long long)sm_count() * 8 ? want : (long long)sm_count() * 8);
        if (cfg)
            ddpm_update_logits_fast_kernel<true, 1><<<grid, 256, 0, STREAM(stream)>>>(x, CBF(logits), CBF(logits_uncond), ldv, cfg_w, modality, seed, offset, t, nullptr, mask_index, text_vocab, pred, R, N, V, r_temp, conf);
        else
            ddpm_update_logits_fast_kernel<false, 1><<<grid, 256, 0, STREAM(stream)>>>(x, CBF(logits), nullptr, ldv, nullptr, modality, seed, offset, t, nullptr, mask_index, text_vocab, pred, R, N, V, r_temp, conf);
    } else {
        int widest = V;
        if (text_vocab > 0) widest = text_vocab > V - text_vocab ? text_vocab : V - text_vocab;
        const size_t smem = (size_t)widest * sizeof(float);
        if (smem > 200 * 1024) { fprintf(stderr, "unidisc_b200: vocabulary range too wide for the maskgit draw (%d)\n", widest); return -1; }
        static bool attr[2] = {false, false};
        if (cfg) {
            if (!attr[1]) { UD_CUDA_CHECK(cudaFuncSetAttribute(maskgit_draw_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr[1] = true; }
            maskgit_draw_kernel<true><<<rows_grid(R), 512, smem, STREAM(stream)>>>(x, CBF(logits), CBF(logits_uncond), ldv, cfg_w, modality, e_noise, gumbel, t, r_temp, mask_index, text_vocab, pred, conf, R, N, V);
        } else {
            if (!attr[0]) { UD_CUDA_CHECK(cudaFuncSetAttribute(maskgit_draw_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr[0] = true; }
            maskgit_draw_kernel<false><<<rows_grid(R), 512, smem, STREAM(stream)>>>(x, CBF(logits), nullptr, ldv, nullptr, modality, e_noise, gumbel, t, r_temp, mask_index, text_vocab, pred, conf, R, N, V);
        }
    }
    UD_CUDA_CHECK(cudaGetLastError());
    const size_t smem_sel = (size_t)N * sizeof(double);
    if (smem_sel > 200 * 1024) { fprintf(stderr, "unidisc_b200: sequence too long for the maskgit selection (%d)\n", N); return -1; }
    static bool attr_sel = false;
    if (smem_sel > 48 * 1024 && !attr_sel) { UD_CUDA_CHECK(cudaFuncSetAttribute(maskgit_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); attr_sel = true; }
    maskgit_select_kernel<<<B, 1024, smem_sel, STREAM(stream)>>>(x, pred, conf, num_unmask, mask_index, out, N);
    UD_CUDA_CHECK(cudaGetLastError());
    return 0;
}

extern "C" int ud_subs_argmax(const void* logits, long long ldv, const int64_t* xt, const int64_t* modality, int64_t* out, int rows,
                              int V, int text_vocab, int mask_index, void* stream) {
    if (rows <= 0) return 0;
    if (ldv % 8 != 0) { fprintf(stderr, "unidisc_b200: subs_argmax needs ldv %% 8 == 0\n"); return -1; }
    subs_argmax_kernel<<<rows_grid(rows), ROW_THREADS, 0, STREAM(stream)>>>(CBF(logits), ldv, xt, modality, out, rows, V, text_vocab, mask_index);
    UD_CUDA_CHECK(cudaGetLastError());
    return 0;
}
