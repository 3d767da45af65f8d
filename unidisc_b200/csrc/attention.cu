// Bidirectional multi-head attention for the DiT block on tcgen05 tensor cores (sm_100a): forward, and a two-kernel
// backward (dK/dV per key tile, dQ per query tile) that recomputes S and dP from Q,K,V,dO and the saved log-sum-exp.
// Replaces torch SDPA / FlexAttention at reference models/dit.py:775-829 (+ document mask model_utils.py:740-771).
//
// Layout: q,k,v,o are 2-D bf16 matrices [B*N, ld] with head h in columns [h*hd, (h+1)*hd) — exactly the packed
// "(three h d)" qkv / "(h d)" output layout of the reference (dit.py:699,846), so no permutes are materialised.
//
// Forward, one CTA per (128-query tile, head, batch), 6 warps:
//   warp0   TMA producer: Q tile once, K/V tiles double-buffered (3-D tensor maps: column, token, batch)
//   warp1   UMMA issuer:  S = Q K^T (SS), O += P V (TS: P is read from TMEM as the A operand, V is the MN-major B operand)
//   warp2-5 softmax: thread r owns query row r: tcgen05.ld S row -> running max / exp2 / row sum -> bf16 P -> tcgen05.st
//           lazy O rescaling (only when the running max grows by > 2^8), final O/l and lse.
// TMEM: S0[128] S1[128] O[hd] P0[64] P1[64] columns.
#include "common.cuh"
#include "unidisc_b200.h"

namespace ud {

static constexpr int ATT_BQ = 128;
static constexpr int ATT_BKV = 128;
static constexpr float LOG2E = 1.4426950408889634f;
static constexpr float LN2 = 0.6931471805599453f;

UD_DEVINL float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct AttnParams {
    int B, N, H;
    float scale_log2;  // softmax scale * log2(e)
    __nv_bfloat16* o;
    long long ldo;
    float* lse;  // [B,H,N]
    const int64_t* sample_ids;  // [B,N] or null
};

template <int HD>
struct AttnSmem {
    static constexpr int TILE_BYTES = 128 * HD * 2;     // one 128-row operand tile
    static constexpr int BOX_BYTES = 128 * 128;          // 128 rows x 64 bf16
    static constexpr int NBOX = HD / 64;
};

// K-major operand tile [128 rows][HD] stored as HD/64 boxes of [128][64]: descriptor for k-step ks (16 elements)
UD_DEVINL uint64_t desc_kmajor(uint32_t tile_base, int ks) {
    return make_smem_desc_sw128(tile_base + (ks >> 2) * (128 * 128) + (ks & 3) * 32, 16, 1024);
}
// MN-major view of the same tile (rows = reduction index, columns = MN): k-step ks covers rows [16ks, 16ks+16)
UD_DEVINL uint64_t desc_mnmajor(uint32_t tile_base, int ks) {
    return make_smem_desc_sw128(tile_base + ks * 2048, 128 * 128, 1024);
}

template <int HD>
__global__ void __launch_bounds__(192, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const AttnParams p) {
    using S = AttnSmem<HD>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + S::TILE_BYTES;            // [2] stages
    uint8_t* sV = sK + 2 * S::TILE_BYTES;        // [2] stages
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * S::TILE_BYTES);
    uint64_t* q_full = bars;           // 1
    uint64_t* k_full = bars + 1;       // 2
    uint64_t* v_full = bars + 3;       // 2
    uint64_t* kv_empty = bars + 5;     // 2
    uint64_t* s_full = bars + 7;       // 2
    uint64_t* p_full = bars + 9;       // 2
    uint64_t* pv_done = bars + 11;     // 1
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 12);
    int* sid_k = reinterpret_cast<int*>(bars + 14);  // [2][128]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * ATT_BQ, h = blockIdx.y, b = blockIdx.z;
    const int T = (p.N + ATT_BKV - 1) / ATT_BKV;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v);
        mbar_init(q_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&k_full[s], 1); mbar_init(&v_full[s], 1); mbar_init(&kv_empty[s], 1);
            mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 128);
        }
        mbar_init(pv_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_ptr_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr_smem;
    const uint32_t tS0 = tmem, tO = tmem + 256, tP0 = tmem + 256 + HD;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(q_full, S::TILE_BYTES);
#pragma unroll
            for (int bx = 0; bx < S::NBOX; ++bx) tma_load_3d(sQ + bx * S::BOX_BYTES, &tm_q, q_full, h * HD + bx * 64, q0, b);
            for (int j = 0; j < T; ++j) {
                const int s = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                mbar_wait(&kv_empty[s], ph ^ 1);
                mbar_expect_tx(&k_full[s], S::TILE_BYTES);
#pragma unroll
                for (int bx = 0; bx < S::NBOX; ++bx)
                    tma_load_3d(sK + s * S::TILE_BYTES + bx * S::BOX_BYTES, &tm_k, &k_full[s], h * HD + bx * 64, j * ATT_BKV, b);
                mbar_expect_tx(&v_full[s], S::TILE_BYTES);
#pragma unroll
                for (int bx = 0; bx < S::NBOX; ++bx)
                    tma_load_3d(sV + s * S::TILE_BYTES + bx * S::BOX_BYTES, &tm_v, &v_full[s], h * HD + bx * 64, j * ATT_BKV, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = make_idesc_bf16(128, ATT_BKV, false, false);
            constexpr uint32_t idesc_pv = make_idesc_bf16(128, HD, false, true);
            const uint32_t aQ = smem_u32(sQ);
            auto issue_s = [&](int j) {
                const int s = j & 1;
                mbar_wait(&k_full[s], (j >> 1) & 1);
                tc_fence_after();
                const uint32_t aK = smem_u32(sK + s * S::TILE_BYTES);
#pragma unroll
                for (int ks = 0; ks < HD / 16; ++ks)
                    umma_ss(tS0 + s * 128, desc_kmajor(aQ, ks), desc_kmajor(aK, ks), idesc_s, ks != 0);
                umma_commit(&s_full[s]);
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            for (int j = 0; j < T; ++j) {
                if (j + 1 < T) issue_s(j + 1);
                const int s = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                mbar_wait(&v_full[s], ph);
                mbar_wait(&p_full[s], ph);
                tc_fence_after();
                const uint32_t aV = smem_u32(sV + s * S::TILE_BYTES);
#pragma unroll
                for (int ks = 0; ks < ATT_BKV / 16; ++ks)
                    umma_ts(tO, tP0 + s * 64 + ks * 8, desc_mnmajor(aV, ks), idesc_pv, (j | ks) != 0);
                umma_commit(&kv_empty[s]);
                umma_commit(pv_done);
            }
        }
    } else {
        // ===================== softmax warps =====================
        const int qd = warp & 3;
        const int rloc = qd * 32 + lane;          // row inside the tile == TMEM lane
        const int row = q0 + rloc;
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
        const bool use_ids = p.sample_ids != nullptr;
        int sid_q = 0;
        if (use_ids) sid_q = row < p.N ? (int)p.sample_ids[(long long)b * p.N + row] : -1;
        const int tid128 = threadIdx.x - 64;
        float m_used = -INFINITY, l = 0.f;
        for (int j = 0; j < T; ++j) {
            const int s = j & 1;
            const uint32_t ph = (j >> 1) & 1;
            if (use_ids) {
                const int kk = j * ATT_BKV + tid128;
                sid_k[s * 128 + tid128] = kk < p.N ? (int)p.sample_ids[(long long)b * p.N + kk] : -2;
                named_bar_sync(1, 128);
            }
            mbar_wait(&s_full[s], ph);
            tc_fence_after();
            float v[128];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tS0 + s * 128 + c * 32 + lane_off, r);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) v[c * 32 + i] = __uint_as_float(r[i]) * p.scale_log2;
            }
            const int kbase = j * ATT_BKV;
            if (use_ids || kbase + ATT_BKV > p.N) {
#pragma unroll
                for (int i = 0; i < 128; ++i) {
                    bool ok = kbase + i < p.N;
                    if (use_ids) ok = ok && (sid_k[s * 128 + i] == sid_q) && (sid_q != -1);
                    if (!ok) v[i] = -INFINITY;
                }
            }
            float mx = v[0];
#pragma unroll
            for (int i = 1; i < 128; ++i) mx = fmaxf(mx, v[i]);
            const float m_new = fmaxf(m_used, mx);
            const bool grow = m_new > m_used + 8.0f;   // also true for -inf -> finite
            if (__any_sync(0xffffffffu, grow)) {
                const float alpha = (m_used == -INFINITY) ? 0.f : ex2(m_used - m_new);
                if (j > 0) {
                    mbar_wait(pv_done, (j - 1) & 1);   // O must be quiescent before it is rescaled
                    tc_fence_after();
#pragma unroll
                    for (int c = 0; c < HD / 32; ++c) {
                        uint32_t r[32];
                        tmem_ld_32x32b_x32(tO + c * 32 + lane_off, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
                        tmem_st_32x32b_x32(tO + c * 32 + lane_off, r);
                    }
                    tmem_st_wait();
                }
                l *= alpha;
                m_used = m_new;
            }
            const float mref = (m_used == -INFINITY) ? 0.f : m_used;
            uint32_t pk_lo[32], pk_hi[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float p0 = ex2(v[2 * i] - mref), p1 = ex2(v[2 * i + 1] - mref);
                const float p2 = ex2(v[64 + 2 * i] - mref), p3 = ex2(v[64 + 2 * i + 1] - mref);
                l += (p0 + p1) + (p2 + p3);
                pk_lo[i] = pack_bf16x2(p0, p1);
                pk_hi[i] = pack_bf16x2(p2, p3);
            }
            tmem_st_32x32b_x32(tP0 + s * 64 + lane_off, pk_lo);
            tmem_st_32x32b_x32(tP0 + s * 64 + 32 + lane_off, pk_hi);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&p_full[s]);
        }
        // ---- epilogue: O / l, lse ----
        mbar_wait(pv_done, (T - 1) & 1);
        tc_fence_after();
        const float inv = l > 0.f ? 1.0f / l : 0.f;
        __nv_bfloat16* orow = p.o + ((long long)b * p.N + row) * p.ldo + h * HD;
#pragma unroll
        for (int c = 0; c < HD / 32; ++c) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(tO + c * 32 + lane_off, r);
            tmem_ld_wait();
            if (row < p.N) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint4 o4;
                    o4.x = pack_bf16x2(__uint_as_float(r[8 * i + 0]) * inv, __uint_as_float(r[8 * i + 1]) * inv);
                    o4.y = pack_bf16x2(__uint_as_float(r[8 * i + 2]) * inv, __uint_as_float(r[8 * i + 3]) * inv);
                    o4.z = pack_bf16x2(__uint_as_float(r[8 * i + 4]) * inv, __uint_as_float(r[8 * i + 5]) * inv);
                    o4.w = pack_bf16x2(__uint_as_float(r[8 * i + 6]) * inv, __uint_as_float(r[8 * i + 7]) * inv);
                    *reinterpret_cast<uint4*>(orow + c * 32 + i * 8) = o4;
                }
            }
        }
        if (row < p.N) p.lse[((long long)b * p.H + h) * p.N + row] = l > 0.f ? (m_used + log2f(l)) * LN2 : INFINITY;
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem);
    }
}

// ------------------------------------------------------------------------------------------------
// delta[b,h,n] = sum_d o[n,d] * do[n,d]   (softmax backward row term)
// ------------------------------------------------------------------------------------------------
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o, long long ldo,
                                  float* __restrict__ delta, int B, int N, int H, int HD) {
    // one warp per (token, head)
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long total = (long long)B * N * H;
    if (gw >= total) return;
    const int hh = (int)(gw % H);
    const long long tok = gw / H;
    const __nv_bfloat16* po = o + tok * ldo + hh * HD;
    const __nv_bfloat16* pd = d_o + tok * ldo + hh * HD;
    float acc = 0.f;
    for (int c = lane * 2; c < HD; c += 64) {
        const uint32_t a = *reinterpret_cast<const uint32_t*>(po + c), g = *reinterpret_cast<const uint32_t*>(pd + c);
        acc += bf16lo(a) * bf16lo(g) + bf16hi(a) * bf16hi(g);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
        const int bb = (int)(tok / N), n = (int)(tok % N);
        delta[((long long)bb * H + hh) * N + n] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// Backward.  MODE 0 (dK,dV): CTA owns a key tile j (rows of the TMEM accumulators = keys) and streams query tiles i:
//     St = K_j Q_i^T, dPt = V_j dO_i^T  ->  Pt = exp2(St*c - lse_q), dSt = Pt*(dPt - delta_q)   (bf16, written over St/dPt)
//     dV += Pt dO_i,  dK += dSt Q_i        (TS MMAs; dO_i / Q_i smem tiles re-read as MN-major B operands)
// MODE 1 (dQ): CTA owns a query tile i (rows = queries) and streams key tiles j:
//     S = Q_i K_j^T,  dP = dO_i V_j^T      ->  dS = P*(dP - delta_row)  (bf16 over dP)
//     dQ += dS K_j                         (K_j re-read as MN-major B)
// "fixed" tiles (K_j,V_j | Q_i,dO_i) are loaded once, "streamed" tiles double-buffered.  The aliasing of the bf16
// probabilities onto the fp32 score columns makes each iteration's MMAs strictly dependent, so the issuer waits for
// its own commits instead of relying on pipelining.
// ------------------------------------------------------------------------------------------------
struct AttnBwdParams {
    int B, N, H;
    float scale_log2, scale;
    const float* lse;
    const float* delta;
    __nv_bfloat16* out0;  // MODE0: dK   MODE1: dQ
    __nv_bfloat16* out1;  // MODE0: dV
    long long ld0, ld1;
    const int64_t* sample_ids;
};

template <int HD, int MODE>
__global__ void __launch_bounds__(192, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tm_fa, const __grid_constant__ CUtensorMap tm_fb,
                const __grid_constant__ CUtensorMap tm_sa, const __grid_constant__ CUtensorMap tm_sb, const AttnBwdParams p) {
    // fixed operands: fa, fb (MODE0: K_j, V_j ; MODE1: Q_i, dO_i); streamed: sa, sb (MODE0: Q_i, dO_i ; MODE1: K_j, V_j)
    using S = AttnSmem<HD>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sFA = smem;
    uint8_t* sFB = sFA + S::TILE_BYTES;
    uint8_t* sSA = sFB + S::TILE_BYTES;      // [2]
    uint8_t* sSB = sSA + 2 * S::TILE_BYTES;  // [2]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sSB + 2 * S::TILE_BYTES);
    uint64_t* f_full = bars;          // 1
    uint64_t* st_full = bars + 1;     // 2
    uint64_t* st_empty = bars + 3;    // 2
    uint64_t* sc_full = bars + 5;     // 1  scores ready (commit)
    uint64_t* pr_full = bars + 6;     // 1  probabilities written (128 arrivals)
    uint64_t* acc_done = bars + 7;    // 1  accumulating MMAs retired (commit)
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 8);
    float* s_lse = reinterpret_cast<float*>(bars + 10);   // [2][128]  (MODE0: per streamed query)
    float* s_dlt = s_lse + 256;                           // [2][128]
    int* s_sid = reinterpret_cast<int*>(s_dlt + 256);     // [2][128]  sample id of the streamed rows

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
    const int T = (p.N + 127) / 128;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_fa); tma_prefetch_desc(&tm_fb); tma_prefetch_desc(&tm_sa); tma_prefetch_desc(&tm_sb);
        mbar_init(f_full, 1);
        for (int s = 0; s < 2; ++s) { mbar_init(&st_full[s], 1); mbar_init(&st_empty[s], 1); }
        mbar_init(sc_full, 1); mbar_init(pr_full, 128); mbar_init(acc_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_ptr_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr_smem;
    const uint32_t tSc = tmem, tDp = tmem + 128, tAcc0 = tmem + 256, tAcc1 = tmem + 256 + HD;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(f_full, 2 * S::TILE_BYTES);
#pragma unroll
            for (int bx = 0; bx < S::NBOX; ++bx) {
                tma_load_3d(sFA + bx * S::BOX_BYTES, &tm_fa, f_full, h * HD + bx * 64, t0, b);
                tma_load_3d(sFB + bx * S::BOX_BYTES, &tm_fb, f_full, h * HD + bx * 64, t0, b);
            }
            for (int i = 0; i < T; ++i) {
                const int s = i & 1;
                mbar_wait(&st_empty[s], ((i >> 1) & 1) ^ 1);
                mbar_expect_tx(&st_full[s], 2 * S::TILE_BYTES);
#pragma unroll
                for (int bx = 0; bx < S::NBOX; ++bx) {
                    tma_load_3d(sSA + s * S::TILE_BYTES + bx * S::BOX_BYTES, &tm_sa, &st_full[s], h * HD + bx * 64, i * 128, b);
                    tma_load_3d(sSB + s * S::TILE_BYTES + bx * S::BOX_BYTES, &tm_sb, &st_full[s], h * HD + bx * 64, i * 128, b);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_sc = make_idesc_bf16(128, 128, false, false);
            constexpr uint32_t idesc_acc = make_idesc_bf16(128, HD, false, true);
            const uint32_t aFA = smem_u32(sFA), aFB = smem_u32(sFB);
            mbar_wait(f_full, 0);
            for (int i = 0; i < T; ++i) {
                const int s = i & 1;
                mbar_wait(&st_full[s], (i >> 1) & 1);
                if (i > 0) mbar_wait(acc_done, (i - 1) & 1);   // previous probabilities (aliased columns) fully consumed
                tc_fence_after();
                const uint32_t aSA = smem_u32(sSA + s * S::TILE_BYTES), aSB = smem_u32(sSB + s * S::TILE_BYTES);
                // scores: Sc = FA . SA^T ; Dp = FB . SB^T     (all K-major, reduction over head_dim)
#pragma unroll
                for (int ks = 0; ks < HD / 16; ++ks) umma_ss(tSc, desc_kmajor(aFA, ks), desc_kmajor(aSA, ks), idesc_sc, ks != 0);
#pragma unroll
                for (int ks = 0; ks < HD / 16; ++ks) umma_ss(tDp, desc_kmajor(aFB, ks), desc_kmajor(aSB, ks), idesc_sc, ks != 0);
                umma_commit(sc_full);
                mbar_wait(pr_full, i & 1);
                tc_fence_after();
                if (MODE == 0) {
                    // dV += Pt . dO_i (B = SB MN-major) ; dK += dSt . Q_i (B = SA MN-major)
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) umma_ts(tAcc1, tSc + ks * 8, desc_mnmajor(aSB, ks), idesc_acc, (i | ks) != 0);
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) umma_ts(tAcc0, tDp + ks * 8, desc_mnmajor(aSA, ks), idesc_acc, (i | ks) != 0);
                } else {
                    // dQ += dS . K_j (B = SA MN-major)
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) umma_ts(tAcc0, tDp + ks * 8, desc_mnmajor(aSA, ks), idesc_acc, (i | ks) != 0);
                }
                umma_commit(&st_empty[s]);
                umma_commit(acc_done);
            }
        }
    } else {
        const int qd = warp & 3;
        const int rloc = qd * 32 + lane;
        const int row = t0 + rloc;   // MODE0: key index ; MODE1: query index
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
        const int tid128 = threadIdx.x - 64;
        const bool use_ids = p.sample_ids != nullptr;
        const long long bh = (long long)b * p.H + h;
        int sid_row = 0;
        if (use_ids) sid_row = row < p.N ? (int)p.sample_ids[(long long)b * p.N + row] : -1;
        float lse_row = 0.f, dlt_row = 0.f;
        if (MODE == 1 && row < p.N) { lse_row = p.lse[bh * p.N + row] * LOG2E; dlt_row = p.delta[bh * p.N + row]; }
        for (int i = 0; i < T; ++i) {
            const int s = i & 1;
            // per-column metadata of the streamed tile
            {
                const int cidx = i * 128 + tid128;
                if (MODE == 0) {
                    s_lse[s * 128 + tid128] = cidx < p.N ? p.lse[bh * p.N + cidx] * LOG2E : INFINITY;
                    s_dlt[s * 128 + tid128] = cidx < p.N ? p.delta[bh * p.N + cidx] : 0.f;
                }
                if (use_ids) s_sid[s * 128 + tid128] = cidx < p.N ? (int)p.sample_ids[(long long)b * p.N + cidx] : -2;
                named_bar_sync(1, 128);
            }
            mbar_wait(sc_full, i & 1);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t rs[32], rd[32];
                tmem_ld_32x32b_x32(tSc + c * 32 + lane_off, rs);
                tmem_ld_32x32b_x32(tDp + c * 32 + lane_off, rd);
                tmem_ld_wait();
                uint32_t pp[16], dd[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    float pv[2], dv[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int col = c * 32 + 2 * e + u;
                        const int cidx = i * 128 + col;
                        bool ok;
                        float lse2, dl;
                        if (MODE == 0) {
                            ok = (cidx < p.N) && (row < p.N);
                            if (use_ids) { const int sq = s_sid[s * 128 + col]; ok = ok && sq == sid_row && sq != -1; }
                            lse2 = s_lse[s * 128 + col]; dl = s_dlt[s * 128 + col];
                        } else {
                            ok = (cidx < p.N) && (row < p.N);
                            if (use_ids) ok = ok && s_sid[s * 128 + col] == sid_row && sid_row != -1;
                            lse2 = lse_row; dl = dlt_row;
                        }
                        const float sc = __uint_as_float(rs[2 * e + u]) * p.scale_log2;
                        const float pr = ok ? ex2(sc - lse2) : 0.f;
                        pv[u] = pr;
                        dv[u] = pr * (__uint_as_float(rd[2 * e + u]) - dl);
                    }
                    pp[e] = pack_bf16x2(pv[0], pv[1]);
                    dd[e] = pack_bf16x2(dv[0], dv[1]);
                }
                if (MODE == 0) tmem_st_32x32b_x16(tSc + c * 16 + lane_off, pp);
                tmem_st_32x32b_x16(tDp + c * 16 + lane_off, dd);
            }
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(pr_full);
        }
        // ---- write the accumulators ----
        mbar_wait(acc_done, (T - 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int a = 0; a < (MODE == 0 ? 2 : 1); ++a) {
            const uint32_t tA = a == 0 ? tAcc0 : tAcc1;
            __nv_bfloat16* dst = (a == 0 ? p.out0 : p.out1) + ((long long)b * p.N + row) * (a == 0 ? p.ld0 : p.ld1) + h * HD;
            const float sc = a == 0 ? p.scale : 1.0f;   // dK, dQ carry the softmax scale; dV does not
#pragma unroll
            for (int c = 0; c < HD / 32; ++c) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tA + c * 32 + lane_off, r);
                tmem_ld_wait();
                if (row < p.N) {
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        uint4 o4;
                        o4.x = pack_bf16x2(__uint_as_float(r[8 * q4 + 0]) * sc, __uint_as_float(r[8 * q4 + 1]) * sc);
                        o4.y = pack_bf16x2(__uint_as_float(r[8 * q4 + 2]) * sc, __uint_as_float(r[8 * q4 + 3]) * sc);
                        o4.z = pack_bf16x2(__uint_as_float(r[8 * q4 + 4]) * sc, __uint_as_float(r[8 * q4 + 5]) * sc);
                        o4.w = pack_bf16x2(__uint_as_float(r[8 * q4 + 6]) * sc, __uint_as_float(r[8 * q4 + 7]) * sc);
                        *reinterpret_cast<uint4*>(dst + c * 32 + q4 * 8) = o4;
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem);
    }
}

template <int HD>
static int attn_smem_bytes(int ntiles) { return ntiles * AttnSmem<HD>::TILE_BYTES + 1024 + 4096; }

static int make_head_tmap(CUtensorMap* tm, const void* base, long long ld, int B, int N, int D) {
    // dims: (column, token, batch); box = 64 columns x 128 tokens x 1 batch
    return make_tmap_3d_bf16(tm, base, (uint64_t)D, (uint64_t)N, (uint64_t)B, (uint64_t)ld, (uint64_t)ld * N, 64, 128, 1);
}

template <int HD>
static int launch_attn_fwd(const void* q, const void* k, long long ldqk, const void* v, long long ldv, const AttnParams& p,
                           cudaStream_t stream) {
    CUtensorMap tq, tk, tv;
    const int D = p.H * HD;
    int rc = make_head_tmap(&tq, q, ldqk, p.B, p.N, D);
    if (rc) return rc;
    if ((rc = make_head_tmap(&tk, k, ldqk, p.B, p.N, D))) return rc;
    if ((rc = make_head_tmap(&tv, v, ldv, p.B, p.N, D))) return rc;
    const int smem = attn_smem_bytes<HD>(5);
    static bool attr = false;
    if (!attr) { UD_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr = true; }
    dim3 grid((p.N + ATT_BQ - 1) / ATT_BQ, p.H, p.B);
    attn_fwd_kernel<HD><<<grid, 192, smem, stream>>>(tq, tk, tv, p);
    UD_CUDA_CHECK(cudaGetLastError());
    return 0;
}

template <int HD>
static int launch_attn_bwd(const void* q, const void* k, long long ldqk, const void* v, long long ldv, const void* d_o,
                           long long ldo, AttnBwdParams p, __nv_bfloat16* dq, __nv_bfloat16* dk, long long lddqk,
                           __nv_bfloat16* dv, long long lddv, cudaStream_t stream) {
    CUtensorMap tq, tk, tv, tdo;
    const int D = p.H * HD;
    int rc;
    if ((rc = make_head_tmap(&tq, q, ldqk, p.B, p.N, D))) return rc;
    if ((rc = make_head_tmap(&tk, k, ldqk, p.B, p.N, D))) return rc;
    if ((rc = make_head_tmap(&tv, v, ldv, p.B, p.N, D))) return rc;
    if ((rc = make_head_tmap(&tdo, d_o, ldo, p.B, p.N, D))) return rc;
    const int smem = attn_smem_bytes<HD>(6);
    static bool attr = false;
    if (!attr) {
        UD_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_kernel<HD, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        UD_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_kernel<HD, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr = true;
    }
    dim3 grid((p.N + 127) / 128, p.H, p.B);
    AttnBwdParams p0 = p;
    p0.out0 = dk; p0.ld0 = lddqk; p0.out1 = dv; p0.ld1 = lddv;
    attn_bwd_kernel<HD, 0><<<grid, 192, smem, stream>>>(tk, tv, tq, tdo, p0);
    UD_CUDA_CHECK(cudaGetLastError());
    AttnBwdParams p1 = p;
    p1.out0 = dq; p1.ld0 = lddqk; p1.out1 = nullptr; p1.ld1 = 0;
    attn_bwd_kernel<HD, 1><<<grid, 192, smem, stream>>>(tq, tdo, tk, tv, p1);
    UD_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace ud

using namespace ud;

extern "C" int ud_attn_fwd(const void* q, const void* k, long long ldqk, const void* v, long long ldv, void* o, long long ldo,
                           float* lse, const int64_t* sample_ids, int B, int N, int H, int head_dim, float scale, void* stream) {
    if (B <= 0 || N <= 0) return 0;
    AttnParams p;
    p.B = B; p.N = N; p.H = H;
    p.scale_log2 = scale * LOG2E;
    p.o = reinterpret_cast<__nv_bfloat16*>(o);
    p.ldo = ldo;
    p.lse = lse;
    p.sample_ids = sample_ids;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (head_dim == 128) return launch_attn_fwd<128>(q, k, ldqk, v, ldv, p, s);
    if (head_dim == 64) return launch_attn_fwd<64>(q, k, ldqk, v, ldv, p, s);
    fprintf(stderr, "unidisc_b200: attention supports head_dim 64 and 128 (got %d)\n", head_dim);
    return -1;
}

extern "C" int ud_attn_bwd(const void* q, const void* k, long long ldqk, const void* v, long long ldv, const void* o,
                           const void* d_o, long long ldo, const float* lse, float* delta, void* dq, void* dk, long long lddqk,
                           void* dv, long long lddv, const int64_t* sample_ids, int B, int N, int H, int head_dim, float scale,
                           void* stream) {
    if (B <= 0 || N <= 0) return 0;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    {
        const long long warps = (long long)B * N * H;
        const int threads = 256;
        const long long blocks = (warps * 32 + threads - 1) / threads;
        attn_delta_kernel<<<(unsigned)blocks, threads, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(o),
                                                                reinterpret_cast<const __nv_bfloat16*>(d_o), ldo, delta, B, N, H, head_dim);
        UD_CUDA_CHECK(cudaGetLastError());
    }
    AttnBwdParams p;
    p.B = B; p.N = N; p.H = H;
    p.scale = scale;
    p.scale_log2 = scale * LOG2E;
    p.lse = lse;
    p.delta = delta;
    p.sample_ids = sample_ids;
    p.out0 = p.out1 = nullptr;
    p.ld0 = p.ld1 = 0;
    auto* dqp = reinterpret_cast<__nv_bfloat16*>(dq);
    auto* dkp = reinterpret_cast<__nv_bfloat16*>(dk);
    auto* dvp = reinterpret_cast<__nv_bfloat16*>(dv);
    if (head_dim == 128) return launch_attn_bwd<128>(q, k, ldqk, v, ldv, d_o, ldo, p, dqp, dkp, lddqk, dvp, lddv, s);
    if (head_dim == 64) return launch_attn_bwd<64>(q, k, ldqk, v, ldv, d_o, ldo, p, dqp, dkp, lddqk, dvp, lddv, s);
    fprintf(stderr, "unidisc_b200: attention supports head_dim 64 and 128 (got %d)\n", head_dim);
    return -1;
}
