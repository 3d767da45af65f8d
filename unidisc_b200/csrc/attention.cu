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
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "unidisc_b200.h"

// ------------------------------------------------------------------------------------------------
// Debug only (-DUD_ATTN_TRACE, tools/attn_trace.py builds its own library with it; the product build never defines it):
// one CTA writes clock64() stamps of its pipeline events into a global table [event][iteration].
// ------------------------------------------------------------------------------------------------
#ifdef UD_ATTN_TRACE
static long long* g_attn_trace_host = nullptr;           // handed to the kernels through their parameter block (no load on the stamp path)
extern "C" int ud_attn_set_trace(void* buf) { g_attn_trace_host = static_cast<long long*>(buf); return 0; }
// UD_TR_INIT evaluates the CTA test once (special-register reads cost tens of cycles each: not on the stamp path)
#define UD_TR_INIT const bool ud_tr_on = (blockIdx.x == 3 && blockIdx.y == 5 && blockIdx.z == 2)
#define UD_TR(cond, ev, it) do { if (ud_tr_on && (cond) && (it) < 64) p.trace[(ev) * 64 + (it)] = clock64(); } while (0)
#else
#define UD_TR_INIT do { } while (0)
#define UD_TR(cond, ev, it) do { } while (0)
#endif

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
// Packed fp32x2 arithmetic (FFMA2 / FADD2 / FMUL2 on sm_100): one instruction for two adjacent scores.  The softmax loops of
// all attention kernels are bound by instruction issue (tools/attn_trace.py), not by a pipe, so halving their FMA-pipe
// instruction count is worth more than any per-pipe balancing.
UD_DEVINL float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
UD_DEVINL uint64_t f2pack(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
UD_DEVINL void f2unpack(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
UD_DEVINL uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
UD_DEVINL uint64_t fmul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
UD_DEVINL uint64_t fsub2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
UD_DEVINL uint64_t fadd2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// (Measured and rejected: computing every 4th exponential with a cubic polynomial on the FMA pipe, the FlashAttention-4 trick.
// MUFU.EX2 is not the limiter of these softmax loops, instruction issue is: the 8 extra FMA/ALU instructions per offloaded
// exponential made the forward 6 % and the backward 1-4 % slower.)

struct AttnParams {
    int B, N, H;       // N = number of QUERY tokens per sample
    int Nk;            // number of key/value tokens per sample (== N except for partial-query attention against a K/V cache)
    float scale_log2;  // softmax scale * log2(e)
    __nv_bfloat16* o;
    long long ldo;
    long long o_bs;    // elements between the first output rows of consecutive samples (N * ldo for a dense [B*N, ldo] matrix)
    float* lse;  // [B,H,N]
    const int64_t* sample_ids;  // [B,N] or null (needs Nk == N)
};

// ------------------------------------------------------------------------------------------------
// Document-mask tile lists.  The reference's FlexAttention BlockMask (model_utils.py:740-771: same sample id, id != -1)
// SKIPS (q-block, kv-block) pairs that cannot match, so a packed batch costs sum(len_i^2), not N^2.  Here every CTA derives
// in its prologue, from the sample ids of its batch row (N int64, L2-resident), the list of streamed tiles whose valid-id
// range [min, max] overlaps the range of its fixed 128-row tile; tiles outside are neither loaded nor multiplied.  A pair of
// tiles that both lie inside ONE document (uniform, equal ids, no padding) is flagged mask-free and takes the unmasked
// softmax instantiation.
// ------------------------------------------------------------------------------------------------
static constexpr int MAX_DOC_TILES = 512;
template <int CAP>
struct DocTilesT {
    int n;
    uint16_t idx[CAP];
    uint8_t nomask[CAP];
    uint8_t code[CAP];
};
using DocTiles = DocTilesT<MAX_DOC_TILES>;

template <int SUB_ROWS, class TL>
UD_DEVINL void doc_tile_list(const int64_t* __restrict__ ids, int N, int f0, int nsub, TL& tl) {   // whole CTA
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    int fmn = 0x7fffffff, fmx = -1;
    bool funi = true;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = f0 + k * 32 + lane;
        const int v = r < N ? (int)ids[r] : -1;
        if (v >= 0) { fmn = min(fmn, v); fmx = max(fmx, v); } else funi = false;
    }
    fmn = __reduce_min_sync(0xffffffffu, fmn);
    fmx = __reduce_max_sync(0xffffffffu, fmx);
    funi = __all_sync(0xffffffffu, funi) && fmn == fmx;
    for (int t = warp; t < nsub; t += nwarps) {
        int smn = 0x7fffffff, smx = -1;
        bool suni = true;
#pragma unroll
        for (int k = 0; k < SUB_ROWS / 32; ++k) {
            const int r = t * SUB_ROWS + k * 32 + lane;
            const int v = r < N ? (int)ids[r] : -1;
            if (v >= 0) { smn = min(smn, v); smx = max(smx, v); } else suni = false;
        }
        smn = __reduce_min_sync(0xffffffffu, smn);
        smx = __reduce_max_sync(0xffffffffu, smx);
        suni = __all_sync(0xffffffffu, suni) && smn == smx;
        const bool act = smx >= fmn && smn <= fmx;          // (a side without any valid id never overlaps)
        if (lane == 0) tl.code[t] = act ? ((funi && suni && smn == fmn) ? 2 : 1) : 0;
    }
    __syncthreads();
    if (warp == 0) {
        int cnt = 0;
        for (int base = 0; base < nsub; base += 32) {
            const int t = base + lane;
            const int c = t < nsub ? tl.code[t] : 0;
            const uint32_t m = __ballot_sync(0xffffffffu, c != 0);
            if (c != 0) {
                const int pos = cnt + __popc(m & ((1u << lane) - 1u));
                tl.idx[pos] = (uint16_t)t;
                tl.nomask[pos] = c == 2;
            }
            cnt += __popc(m);
        }
        if (lane == 0) tl.n = cnt;
    }
    __syncthreads();
}

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
UD_DEVINL uint64_t desc_kmajor64(uint32_t tile_base, int ks) {   // [64 rows][HD] as HD/64 boxes of [64][64]
    return make_smem_desc_sw128(tile_base + (ks >> 2) * (64 * 128) + (ks & 3) * 32, 16, 1024);
}
UD_DEVINL uint64_t desc_mnmajor64(uint32_t tile_base, int ks) {  // MN-major view: k-step = 16 rows, MN chunks 8 KB apart
    return make_smem_desc_sw128(tile_base + ks * 2048, 64 * 128, 1024);
}

// ------------------------------------------------------------------------------------------------
// Forward v3: the v1 dataflow (one 128-query tile per CTA, double-buffered S and P in TMEM so QK^T of tile j+1 runs under
// the softmax of tile j) with EIGHT softmax warps: warpgroup g owns key columns [64g, 64g+64) of every S tile, i.e. each
// thread handles 64 scores (two tcgen05.ld in flight, one wait), the two halves exchange their row maxima / sums through
// shared memory, and every SM sub-partition has two softmax warps to overlap MUFU, ALU and TMEM latency.
// ------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(320, 1)
attn_fwd3_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
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
    uint64_t* v_empty = bars + 5;      // 2   V stage free once P.V of that tile retired
    uint64_t* s_full = bars + 7;       // 2
    uint64_t* p_full = bars + 9;       // 2
    uint64_t* pv_done = bars + 11;     // 1
    uint64_t* k_empty = bars + 12;     // 2   K stage free as soon as Q.K^T of that tile retired (long before P.V):
                                       //     the next K tile is prefetched under the softmax instead of after it
    uint64_t* all_done = bars + 14;    // 1   every MMA of this CTA has retired (completes exactly ONCE: the epilogue's wait cannot
                                       //     alias an earlier phase the way a parity wait on the per-tile pv_done barrier can
                                       //     when two completions are still outstanding)
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 15);
    __shared__ int sid_k[2 * 128];            // [stage][128]
    __shared__ float xch[2 * 2 * 128];        // [stage][warpgroup][row]: partial row maxima (and final row sums)
    __shared__ DocTiles tl;                   // active key tiles of this query tile (document mask only)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * ATT_BQ, h = blockIdx.y, b = blockIdx.z;
    const int Tall = (p.Nk + ATT_BKV - 1) / ATT_BKV;
    const bool use_ids = p.sample_ids != nullptr;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v);
        mbar_init(q_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&k_full[s], 1); mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1); mbar_init(&k_empty[s], 1);
            mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 256);
        }
        mbar_init(pv_done, 1);
        mbar_init(all_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_ptr_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr_smem;
    const uint32_t tS0 = tmem, tO = tmem + 256, tP0 = tmem + 256 + HD;
    if (use_ids) doc_tile_list<ATT_BKV>(p.sample_ids + (long long)b * p.N, p.N, q0, Tall, tl);
    // T = number of key tiles this CTA visits; iteration jj works on key tile tile_of(jj) (stage / phase bookkeeping is in jj)
    const int T = use_ids ? tl.n : Tall;
    auto tile_of = [&](int jj) { return use_ids ? (int)tl.idx[jj] : jj; };

    if (warp == 0) {
        if (lane == 0 && T > 0) {
            mbar_expect_tx(q_full, S::TILE_BYTES);
#pragma unroll
            for (int bx = 0; bx < S::NBOX; ++bx) tma_load_3d(sQ + bx * S::BOX_BYTES, &tm_q, q_full, h * HD + bx * 64, q0, b);
            // K(jj+1) is requested BEFORE V(jj): its stage is free as soon as QK^T(jj-1) retired, whereas V(jj) has to wait for
            // P.V(jj-2) — issued in program order behind V, the K tile would arrive a whole tile later than it could
            auto load_k = [&](int jj) {
                const int j = tile_of(jj);
                const int s = jj & 1;
                mbar_wait(&k_empty[s], ((jj >> 1) & 1) ^ 1);
                mbar_expect_tx(&k_full[s], S::TILE_BYTES);
#pragma unroll
                for (int bx = 0; bx < S::NBOX; ++bx)
                    tma_load_3d(sK + s * S::TILE_BYTES + bx * S::BOX_BYTES, &tm_k, &k_full[s], h * HD + bx * 64, j * ATT_BKV, b);
            };
            load_k(0);
            for (int jj = 0; jj < T; ++jj) {
                const int j = tile_of(jj);
                const int s = jj & 1;
                const uint32_t ph = (jj >> 1) & 1;
                if (jj + 1 < T) load_k(jj + 1);
                mbar_wait(&v_empty[s], ph ^ 1);
                mbar_expect_tx(&v_full[s], S::TILE_BYTES);
#pragma unroll
                for (int bx = 0; bx < S::NBOX; ++bx)
                    tma_load_3d(sV + s * S::TILE_BYTES + bx * S::BOX_BYTES, &tm_v, &v_full[s], h * HD + bx * 64, j * ATT_BKV, b);
            }
        }
    } else if (warp == 1) {
        if (T > 0) {
            // warp-uniform issue loop, one elected lane executes the tcgen05 instructions (see elect_one)
            const uint32_t leader = elect_one();
            constexpr uint32_t idesc_s = make_idesc_bf16(128, ATT_BKV, false, false);
            constexpr uint32_t idesc_pv = make_idesc_bf16(128, HD, false, true);
            const uint32_t aQ = smem_u32(sQ);
            auto issue_s = [&](int j) {
                const int s = j & 1;
                mbar_wait(&k_full[s], (j >> 1) & 1);
                tc_fence_after();
                const uint32_t aK = smem_u32(sK + s * S::TILE_BYTES);
                if (leader) {
#pragma unroll
                    for (int ks = 0; ks < HD / 16; ++ks)
                        umma_ss(tS0 + s * 128, desc_kmajor(aQ, ks), desc_kmajor(aK, ks), idesc_s, ks != 0);
                    umma_commit(&s_full[s]);
                    umma_commit(&k_empty[s]);
                }
                __syncwarp();
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
                const uint32_t acc0 = j != 0;
                if (leader) {
#pragma unroll
                    for (int ks = 0; ks < ATT_BKV / 16; ++ks)
                        umma_ts(tO, tP0 + s * 64 + ks * 8, desc_mnmajor(aV, ks), idesc_pv, acc0 | (ks != 0));
                    umma_commit(&v_empty[s]);
                    umma_commit(pv_done);
                    if (j + 1 == T) umma_commit(all_done);
                }
                __syncwarp();
            }
        }
    } else {
        const int wg = (warp - 2) >> 2;           // which half of the key columns
        const int qd = warp & 3;
        const int rloc = qd * 32 + lane;          // row inside the tile == TMEM lane
        const int row = q0 + rloc;
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
        int sid_q = 0;
        if (use_ids) sid_q = row < p.N ? (int)p.sample_ids[(long long)b * p.N + row] : -1;
        const int tid256 = threadIdx.x - 64;
        const float scl = p.scale_log2;
        const int Ntok = p.Nk;
        float m_used = -INFINITY, l = 0.f;
        for (int jj = 0; jj < T; ++jj) {
            const int j = tile_of(jj);
            const int s = jj & 1;
            const uint32_t ph = (jj >> 1) & 1;
            const bool use_ids = p.sample_ids != nullptr && !tl.nomask[jj];     // this tile pair needs the per-element mask
            if (use_ids && tid256 < 128) {
                const int kk = j * ATT_BKV + tid256;
                sid_k[s * 128 + tid256] = kk < Ntok ? (int)p.sample_ids[(long long)b * Ntok + kk] : -2;
            }
            mbar_wait(&s_full[s], ph);
            tc_fence_after();
            uint32_t r0[32], r1[32];
            tmem_ld_32x32b_x32(tS0 + s * 128 + wg * 64 + lane_off, r0);
            tmem_ld_32x32b_x32(tS0 + s * 128 + wg * 64 + 32 + lane_off, r1);
            tmem_ld_wait();
            const int kbase = j * ATT_BKV + wg * 64;
            const bool tail = kbase + 64 > Ntok;
            if (tail && !use_ids) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (kbase + i >= Ntok) r0[i] = 0xff800000u;
                    if (kbase + 32 + i >= Ntok) r1[i] = 0xff800000u;
                }
            }
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
            if (!use_ids) {
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    mx0 = fmaxf(mx0, __uint_as_float(r0[i])); mx1 = fmaxf(mx1, __uint_as_float(r0[i + 1]));
                    mx2 = fmaxf(mx2, __uint_as_float(r1[i])); mx3 = fmaxf(mx3, __uint_as_float(r1[i + 1]));
                }
            }
            float* xs = xch + s * 256;
            if (!use_ids) xs[wg * 128 + rloc] = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
            named_bar_sync(1, 256);                       // also publishes sid_k for this tile
            if (use_ids) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const bool ok0 = kbase + i < Ntok && sid_k[s * 128 + wg * 64 + i] == sid_q && sid_q != -1;
                    const bool ok1 = kbase + 32 + i < Ntok && sid_k[s * 128 + wg * 64 + 32 + i] == sid_q && sid_q != -1;
                    if (!ok0) r0[i] = 0xff800000u;
                    if (!ok1) r1[i] = 0xff800000u;
                    mx0 = fmaxf(mx0, __uint_as_float(r0[i])); mx2 = fmaxf(mx2, __uint_as_float(r1[i]));
                }
                xs[wg * 128 + rloc] = fmaxf(mx0, mx2);
                named_bar_sync(2, 256);
            }
            const float mx = fmaxf(xs[rloc], xs[128 + rloc]) * scl;
            const float m_new = fmaxf(m_used, mx);
            const bool grow = m_new > m_used + 8.0f;      // identical decision in both warpgroups (same inputs)
            if (__any_sync(0xffffffffu, grow)) {
                const float alpha = (m_used == -INFINITY) ? 0.f : ex2(m_used - m_new);
                if (jj > 0) {
                    mbar_wait(pv_done, (jj - 1) & 1);      // O must be quiescent before it is rescaled
                    tc_fence_after();
#pragma unroll
                    for (int c = 0; c < HD / 64; ++c) {   // each warpgroup rescales its half of O's columns
                        uint32_t r[32];
                        tmem_ld_32x32b_x32(tO + wg * (HD / 2) + c * 32 + lane_off, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
                        tmem_st_32x32b_x32(tO + wg * (HD / 2) + c * 32 + lane_off, r);
                    }
                    tmem_st_wait();
                }
                l *= alpha;
                m_used = m_new;
            }
            const float mref = (m_used == -INFINITY) ? 0.f : m_used;
            float l0 = 0.f, l1 = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float a0 = ex2(fmaf(__uint_as_float(r0[2 * i]), scl, -mref)), a1 = ex2(fmaf(__uint_as_float(r0[2 * i + 1]), scl, -mref));
                l0 += a0 + a1;
                r0[i] = pack_bf16x2(a0, a1);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float b0 = ex2(fmaf(__uint_as_float(r1[2 * i]), scl, -mref)), b1 = ex2(fmaf(__uint_as_float(r1[2 * i + 1]), scl, -mref));
                l1 += b0 + b1;
                r0[16 + i] = pack_bf16x2(b0, b1);
            }
            l += l0 + l1;
            tmem_st_32x32b_x32(tP0 + s * 64 + wg * 32 + lane_off, r0);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&p_full[s]);
        }
        // ---- epilogue: combine the two half-row sums, O / l, lse ----
        float* xs = xch;                                   // last tile used xch[(T-1)&1]; use the other slot for the sums
        xs += ((T & 1) ? 256 : 0);
        xs[wg * 128 + rloc] = l;
        named_bar_sync(1, 256);
        const float lt = xs[rloc] + xs[128 + rloc];
        if (T > 0) mbar_wait(all_done, 0);
        tc_fence_after();
        const float inv = lt > 0.f ? 1.0f / lt : 0.f;
        __nv_bfloat16* orow = p.o + (long long)b * p.o_bs + (long long)row * p.ldo + h * HD + wg * (HD / 2);
#pragma unroll
        for (int c = 0; c < HD / 64; ++c) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(tO + wg * (HD / 2) + c * 32 + lane_off, r);
            tmem_ld_wait();
            if (inv == 0.f) {                              // fully masked row (padding): O was never written / holds no mass
#pragma unroll
                for (int i = 0; i < 32; ++i) r[i] = 0u;
            }
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
        if (wg == 0 && row < p.N) p.lse[((long long)b * p.H + h) * p.N + row] = lt > 0.f ? (m_used + log2f(lt)) * LN2 : INFINITY;
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem);
    }
}

// ------------------------------------------------------------------------------------------------
// Forward v6: v3's dataflow (double-buffered S so that Q.K^T of tile j+1 runs under the softmax of tile j, and the softmax warps
// never wait for P.V) at HALF the key-tile width, which makes everything small enough for TWO CTAs per SM:
//   key tiles of 64: S0[64] S1[64] O[HD] = 256 TMEM columns, Q + 2 x (K, V) 64-row stages = 96 KB of shared memory (hd = 128);
//   four softmax warps, one thread per query row: the 64 scores of a tile live in registers for ONE pass (no cross-warp
//   exchange of row maxima as in v3); bf16 P overwrites the first 32 columns of its S buffer.
// The softmax warps of a CTA run back to back (their only waits are S(j) — computed a tile ahead — and the rare lazy rescale),
// two CTAs share each SM's MUFU, and the tensor pipe / TMA of one CTA fill the gaps of the other.
// ------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(192, 2)
attn_fwd6_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                 const __grid_constant__ CUtensorMap tm_v, const AttnParams p) {
    constexpr int BKV = 64;
    constexpr int QTILE = 128 * HD * 2;          // bytes of the Q tile
    constexpr int KTILE = BKV * HD * 2;          // bytes of one K (or V) stage
    constexpr int NBOX = HD / 64;
    constexpr int QBOX = 128 * 128, KBOX = BKV * 128;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + QTILE;                    // [2] stages
    uint8_t* sV = sK + 2 * KTILE;                // [2] stages
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * KTILE);
    uint64_t* q_full = bars;           // 1
    uint64_t* k_full = bars + 1;       // 2
    uint64_t* v_full = bars + 3;       // 2
    uint64_t* k_empty = bars + 5;      // 2   K stage free once Q.K^T of that tile retired
    uint64_t* v_empty = bars + 7;      // 2   V stage free once P.V of that tile retired
    uint64_t* s_full = bars + 9;       // 2
    uint64_t* p_full = bars + 11;      // 2   128 arrivals
    uint64_t* pv_done = bars + 13;     // 1   completes once per tile; waited only in step (lazy rescale)
    uint64_t* all_done = bars + 14;    // 1   completes once
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 15);
    __shared__ int sid_k[2 * BKV];            // [stage][64]
    __shared__ DocTiles tl;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * ATT_BQ, h = blockIdx.y, b = blockIdx.z;
    const int Tall = (p.Nk + BKV - 1) / BKV;
    const bool use_ids = p.sample_ids != nullptr;
    constexpr uint32_t TCOLS = (128 + HD <= 256) ? 256 : 512;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v);
        mbar_init(q_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&k_full[s], 1); mbar_init(&v_full[s], 1); mbar_init(&k_empty[s], 1); mbar_init(&v_empty[s], 1);
            mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 128);
        }
        mbar_init(pv_done, 1); mbar_init(all_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<TCOLS>(tmem_ptr_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr_smem;
    const uint32_t tS0 = tmem, tO0 = tmem + 128;        // S_s = tS0 + 64 s
    if (use_ids) doc_tile_list<BKV>(p.sample_ids + (long long)b * p.N, p.N, q0, Tall, tl);
    const int T = use_ids ? tl.n : Tall;
    auto tile_of = [&](int jj) { return use_ids ? (int)tl.idx[jj] : jj; };

    if (warp == 0) {
        if (lane == 0 && T > 0) {
            mbar_expect_tx(q_full, QTILE);
#pragma unroll
            for (int bx = 0; bx < NBOX; ++bx) tma_load_3d(sQ + bx * QBOX, &tm_q, q_full, h * HD + bx * 64, q0, b);
            auto load_k = [&](int jj) {
                const int s = jj & 1;
                mbar_wait(&k_empty[s], ((jj >> 1) & 1) ^ 1);
                mbar_expect_tx(&k_full[s], KTILE);
#pragma unroll
                for (int bx = 0; bx < NBOX; ++bx)
                    tma_load_3d(sK + s * KTILE + bx * KBOX, &tm_k, &k_full[s], h * HD + bx * 64, tile_of(jj) * BKV, b);
            };
            load_k(0);
            for (int jj = 0; jj < T; ++jj) {
                const int s = jj & 1;
                if (jj + 1 < T) load_k(jj + 1);       // K first: its stage frees a tile earlier than the V stage
                mbar_wait(&v_empty[s], ((jj >> 1) & 1) ^ 1);
                mbar_expect_tx(&v_full[s], KTILE);
#pragma unroll
                for (int bx = 0; bx < NBOX; ++bx)
                    tma_load_3d(sV + s * KTILE + bx * KBOX, &tm_v, &v_full[s], h * HD + bx * 64, tile_of(jj) * BKV, b);
            }
        }
    } else if (warp == 1) {
        if (T > 0) {
            const uint32_t leader = elect_one();
            constexpr uint32_t idesc_s = make_idesc_bf16(128, BKV, false, false);
            constexpr uint32_t idesc_pv = make_idesc_bf16(128, HD, false, true);
            const uint32_t aQ = smem_u32(sQ);
            auto issue_s = [&](int j) {
                const int s = j & 1;
                mbar_wait(&k_full[s], (j >> 1) & 1);
                tc_fence_after();
                const uint32_t aK = smem_u32(sK + s * KTILE);
                if (leader) {
#pragma unroll
                    for (int ks = 0; ks < HD / 16; ++ks)
                        umma_ss(tS0 + s * 64, desc_kmajor(aQ, ks), desc_kmajor64(aK, ks), idesc_s, ks != 0);
                    umma_commit(&s_full[s]);
                    umma_commit(&k_empty[s]);
                }
                __syncwarp();
            };
            mbar_wait(q_full, 0);
            issue_s(0);
            for (int j = 0; j < T; ++j) {
                if (j + 1 < T) issue_s(j + 1);        // into the other S buffer (its P was consumed by P.V(j-1), issued earlier)
                const int s = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                mbar_wait(&v_full[s], ph);
                mbar_wait(&p_full[s], ph);
                tc_fence_after();
                const uint32_t aV = smem_u32(sV + s * KTILE);
                if (leader) {
#pragma unroll
                    for (int ks = 0; ks < BKV / 16; ++ks)
                        umma_ts(tO0, tS0 + s * 64 + ks * 8, desc_mnmajor64(aV, ks), idesc_pv, (uint32_t)(j != 0) | (uint32_t)(ks != 0));
                    umma_commit(&v_empty[s]);
                    umma_commit(pv_done);
                    if (j + 1 == T) umma_commit(all_done);
                }
                __syncwarp();
            }
        }
    } else {
        const int qd = warp & 3;
        const int rloc = qd * 32 + lane;          // row inside the tile == TMEM lane
        const int row = q0 + rloc;
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
        const uint32_t tO = tO0 + lane_off;
        const int tid128 = threadIdx.x - 64;
        int sid_q = 0;
        if (use_ids) sid_q = row < p.N ? (int)p.sample_ids[(long long)b * p.N + row] : -1;
        const float scl = p.scale_log2;
        const int Ntok = p.Nk;
        float m_used = -INFINITY, l = 0.f;
        for (int jj = 0; jj < T; ++jj) {
            const int j = tile_of(jj);
            const int s = jj & 1;
            const uint32_t ph = (jj >> 1) & 1;
            const bool use_mask = use_ids && !tl.nomask[jj];
            const int kbase = j * BKV;
            const bool tail = kbase + BKV > Ntok;
            if (use_mask) {
                if (tid128 < BKV) {
                    const int kk = kbase + tid128;
                    sid_k[s * BKV + tid128] = kk < Ntok ? (int)p.sample_ids[(long long)b * Ntok + kk] : -2;
                }
                named_bar_sync(1, 128);
            }
            mbar_wait(&s_full[s], ph);
            tc_fence_after();
            const uint32_t tS = tS0 + s * 64 + lane_off;
            uint32_t r0[32], r1[32];
            tmem_ld_32x32b_x32(tS, r0);
            tmem_ld_32x32b_x32(tS + 32, r1);
            tmem_ld_wait();
            if (use_mask) {
                const int* sk = sid_k + s * BKV;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (!(kbase + i < Ntok && sk[i] == sid_q && sid_q != -1)) r0[i] = 0xff800000u;
                    if (!(kbase + 32 + i < Ntok && sk[32 + i] == sid_q && sid_q != -1)) r1[i] = 0xff800000u;
                }
            } else if (tail) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (kbase + i >= Ntok) r0[i] = 0xff800000u;
                    if (kbase + 32 + i >= Ntok) r1[i] = 0xff800000u;
                }
            }
            // row maximum with 3-input FMNMX3: one instruction per two scores
            float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                mx0 = fmax3(mx0, __uint_as_float(r0[i]), __uint_as_float(r0[i + 1]));
                mx1 = fmax3(mx1, __uint_as_float(r0[i + 2]), __uint_as_float(r0[i + 3]));
                mx2 = fmax3(mx2, __uint_as_float(r1[i]), __uint_as_float(r1[i + 1]));
                mx3 = fmax3(mx3, __uint_as_float(r1[i + 2]), __uint_as_float(r1[i + 3]));
            }
            const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * scl;
            const float m_new = fmaxf(m_used, mx);
            const bool grow = m_new > m_used + 8.0f;
            if (__any_sync(0xffffffffu, grow)) {
                const float alpha = (m_used == -INFINITY) ? 0.f : ex2(m_used - m_new);
                if (jj > 0) {
                    mbar_wait(pv_done, (jj - 1) & 1);          // O must be quiescent before it is rescaled
                    tc_fence_after();
#pragma unroll
                    for (int c = 0; c < HD / 32; ++c) {
                        uint32_t r[32];
                        tmem_ld_32x32b_x32(tO + c * 32, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
                        tmem_st_32x32b_x32(tO + c * 32, r);
                    }
                    tmem_st_wait();
                }
                l *= alpha;
                m_used = m_new;
            }
            const float mref = (m_used == -INFINITY) ? 0.f : m_used;
            // two scores per FFMA2 / FADD2 (the pairs are adjacent TMEM columns = adjacent registers)
            const uint64_t scl2 = f2pack(scl, scl), nm2 = f2pack(-mref, -mref);
            uint64_t l2a = 0ull, l2b = 0ull;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float x0, x1;
                f2unpack(ffma2(f2pack(__uint_as_float(r0[2 * i]), __uint_as_float(r0[2 * i + 1])), scl2, nm2), x0, x1);
                const float a0 = ex2(x0), a1 = ex2(x1);
                l2a = fadd2(l2a, f2pack(a0, a1));
                r0[i] = pack_bf16x2(a0, a1);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float x0, x1;
                f2unpack(ffma2(f2pack(__uint_as_float(r1[2 * i]), __uint_as_float(r1[2 * i + 1])), scl2, nm2), x0, x1);
                const float b0 = ex2(x0), b1 = ex2(x1);
                l2b = fadd2(l2b, f2pack(b0, b1));
                r0[16 + i] = pack_bf16x2(b0, b1);
            }
            {
                float s0, s1;
                f2unpack(fadd2(l2a, l2b), s0, s1);
                l += s0 + s1;
            }
            tmem_st_32x32b_x32(tS, r0);                   // P over the first 32 columns of this S buffer
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&p_full[s]);
        }
        // ---- epilogue: O / l, lse ----
        if (T > 0) mbar_wait(all_done, 0);
        tc_fence_after();
        const float inv = l > 0.f ? 1.0f / l : 0.f;
        __nv_bfloat16* orow = p.o + (long long)b * p.o_bs + (long long)row * p.ldo + h * HD;
#pragma unroll
        for (int c = 0; c < HD / 32; ++c) {
            uint32_t r[32];
            if (T > 0) {
                tmem_ld_32x32b_x32(tO + c * 32, r);
                tmem_ld_wait();
            }
            if (inv == 0.f) {                              // fully masked row (padding): O holds no mass / was never written
#pragma unroll
                for (int i = 0; i < 32; ++i) r[i] = 0u;
            }
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
        tmem_dealloc<TCOLS>(tmem);
    }
}

// ------------------------------------------------------------------------------------------------
// delta[b,h,n] = sum_d o[n,d] * do[n,d]   (softmax backward row term)
// ------------------------------------------------------------------------------------------------
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o, long long ldo,
                                  float* __restrict__ delta, int B, int N, int H, int HD,
                                  const float* __restrict__ lse = nullptr, float* __restrict__ lse2 = nullptr) {
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
        // the v3 dK/dV kernel reads the per-query log-sum-exp already in the exp2 domain (one FFMA + EX2 per score)
        if (lse2 != nullptr) lse2[((long long)bb * H + hh) * N + n] = -lse[((long long)bb * H + hh) * N + n] * 1.4426950408889634f;
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
    int safe_order;
    const float* lse2;    // v3: -lse * log2(e) (added by the FFMA in front of exp2), written by the delta pass (second half of the caller's delta scratch)
    // v2 dQ kernel only: it forms delta = rowsum(dO * O) for its own 128 query rows (and publishes it for the dK/dV kernel,
    // which runs after it), so no separate pass over O and dO is launched
    const __nv_bfloat16* o;
    const __nv_bfloat16* d_o;
    long long ldo;
    float* delta_out;
#ifdef UD_ATTN_TRACE
    long long* trace;
#endif
};

// ------------------------------------------------------------------------------------------------
// Backward v2: same maths as attn_bwd_kernel, but the streamed operand is cut into 64-row sub-tiles and the score
// accumulators are double-buffered in TMEM, so the tensor core (scores of sub-tile i+2, accumulation of sub-tile i)
// runs concurrently with the exp / dS arithmetic of sub-tile i+1 instead of strictly alternating with it.
//   TMEM: ScA[64] DpA[64] ScB[64] DpB[64] | Acc0[HD] Acc1[HD]          smem: 2 fixed 128-row tiles + 4-stage ring of 64-row tiles
// ------------------------------------------------------------------------------------------------
template <int HD>
struct AttnBwd2Smem {
    static constexpr int FIX_BYTES = 128 * HD * 2;
    static constexpr int SUB_BYTES = 64 * HD * 2;       // one streamed 64-row tile
    static constexpr int SUB_BOX = 64 * 128;            // 64 rows x 64 bf16
    static constexpr int NST = 4;
    static constexpr int NBOX = HD / 64;
    static constexpr int BYTES = 2 * FIX_BYTES + NST * 2 * SUB_BYTES + 1024 + 4096;
};

template <int HD, int MODE>
__global__ void __launch_bounds__(320, 1)
attn_bwd2_kernel(const __grid_constant__ CUtensorMap tm_fa, const __grid_constant__ CUtensorMap tm_fb,
                 const __grid_constant__ CUtensorMap tm_sa, const __grid_constant__ CUtensorMap tm_sb, const AttnBwdParams p) {
    using S = AttnBwd2Smem<HD>;
    constexpr int NST = S::NST;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sFA = smem;
    uint8_t* sFB = sFA + S::FIX_BYTES;
    uint8_t* sST = sFB + S::FIX_BYTES;                   // stage s: [SA | SB]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sST + NST * 2 * S::SUB_BYTES);
    uint64_t* f_full = bars;                 // 1
    uint64_t* st_full = bars + 1;            // NST
    uint64_t* st_empty = bars + 1 + NST;     // NST
    uint64_t* s_full = bars + 1 + 2 * NST;   // 2   S^T of a sub-tile in TMEM      (tensor pipe -> softmax warpgroup)
    uint64_t* dp_full = s_full + 2;          // 2   dP^T of the same sub-tile
    uint64_t* p_rdy = dp_full + 2;           // 2   bf16 P^T written over S^T        (softmax warpgroup -> tensor pipe)
    uint64_t* ds_rdy = p_rdy + 2;            // 2   bf16 dS^T written over dP^T
    uint64_t* acc_done = ds_rdy + 2;         // 2
    uint64_t* all_done = acc_done + 2;       // 1   every MMA of this CTA has retired.  Completes exactly ONCE, so the epilogue's wait
                                             //     cannot alias an earlier phase: a parity wait on acc_done[] by the warpgroup that
                                             //     does not own that buffer passed early whenever two completions were still
                                             //     outstanding (rare, timing dependent: wrong dK / dV)
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(all_done + 1);
    __shared__ __align__(16) float s_lse[4 * 64];   // per-column metadata of the streamed sub-tile: [warpgroup][2 slots][64]
    __shared__ __align__(16) float s_dlt[4 * 64];   // (two alternating slots per warpgroup: a fast thread may already publish
    __shared__ __align__(16) int s_sid[4 * 64];     //  sub-tile i+2 while a slow one still reads sub-tile i)
    __shared__ DocTiles tl;                         // active streamed sub-tiles of this fixed tile (document mask only)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
    const int T2all = (p.N + 63) / 64;
    const bool use_ids = p.sample_ids != nullptr;
    UD_TR_INIT;
    UD_TR(threadIdx.x == 0, 17, 0);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_fa); tma_prefetch_desc(&tm_fb); tma_prefetch_desc(&tm_sa); tma_prefetch_desc(&tm_sb);
        mbar_init(f_full, 1);
        for (int s = 0; s < NST; ++s) { mbar_init(&st_full[s], 1); mbar_init(&st_empty[s], 1); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&s_full[s], 1); mbar_init(&dp_full[s], 1); mbar_init(&p_rdy[s], 128); mbar_init(&ds_rdy[s], 128);
            mbar_init(&acc_done[s], 1);
        }
        mbar_init(all_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_ptr_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr_smem;
    const uint32_t tAcc0 = tmem + 256, tAcc1 = tmem + 256 + HD;
    if (use_ids) doc_tile_list<64>(p.sample_ids + (long long)b * p.N, p.N, t0, T2all, tl);
    // T2 = number of streamed 64-row sub-tiles this CTA visits; iteration ii works on sub-tile sub_of(ii) (stage / buffer /
    // phase bookkeeping is in ii, coordinates and per-column metadata in the sub-tile index)
    const int T2 = use_ids ? tl.n : T2all;
    auto sub_of = [&](int ii) { return use_ids ? (int)tl.idx[ii] : ii; };

    if (warp == 0) {
        if (lane == 0 && T2 > 0) {
            mbar_expect_tx(f_full, 2 * S::FIX_BYTES);
#pragma unroll
            for (int bx = 0; bx < S::NBOX; ++bx) {
                tma_load_3d(sFA + bx * (128 * 128), &tm_fa, f_full, h * HD + bx * 64, t0, b);
                tma_load_3d(sFB + bx * (128 * 128), &tm_fb, f_full, h * HD + bx * 64, t0, b);
            }
            for (int i = 0; i < T2; ++i) {
                const int s = i % NST;
                const int sub = sub_of(i);
                mbar_wait(&st_empty[s], ((i / NST) & 1) ^ 1);
                UD_TR(true, 13, i);
                mbar_expect_tx(&st_full[s], 2 * S::SUB_BYTES);
                uint8_t* sa = sST + s * 2 * S::SUB_BYTES;
                uint8_t* sb = sa + S::SUB_BYTES;
#pragma unroll
                for (int bx = 0; bx < S::NBOX; ++bx) {
                    tma_load_3d(sa + bx * S::SUB_BOX, &tm_sa, &st_full[s], h * HD + bx * 64, sub * 64, b);
                    tma_load_3d(sb + bx * S::SUB_BOX, &tm_sb, &st_full[s], h * HD + bx * 64, sub * 64, b);
                }
            }
        }
    } else if (warp == 1) {
        if (T2 > 0) {
            // the whole warp runs this loop (warp-uniform control flow and descriptor arithmetic); one elected lane issues
            const uint32_t leader = elect_one();
            constexpr uint32_t idesc_sc = make_idesc_bf16(128, 64, false, false);
            constexpr uint32_t idesc_acc = make_idesc_bf16(128, HD, false, true);
            const uint32_t aFA = smem_u32(sFA), aFB = smem_u32(sFB);
            auto issue_scores = [&](int i) {
                const int s = i % NST, bf = i & 1;
                mbar_wait(&st_full[s], (i / NST) & 1);
                tc_fence_after();
                UD_TR(leader, 0, i);
                const uint32_t aSA = smem_u32(sST + s * 2 * S::SUB_BYTES), aSB = aSA + S::SUB_BYTES;
                const uint32_t tSc = tmem + bf * 128, tDp = tSc + 64;
                if (leader) {
                    // S^T and dP^T are signalled separately: the warpgroup starts its exponentials while dP^T is still being
                    // multiplied
#pragma unroll
                    for (int ks = 0; ks < HD / 16; ++ks) umma_ss(tSc, desc_kmajor(aFA, ks), desc_kmajor64(aSA, ks), idesc_sc, ks != 0);
                    umma_commit(&s_full[bf]);
#pragma unroll
                    for (int ks = 0; ks < HD / 16; ++ks) umma_ss(tDp, desc_kmajor(aFB, ks), desc_kmajor64(aSB, ks), idesc_sc, ks != 0);
                    umma_commit(&dp_full[bf]);
                }
                UD_TR(leader, 1, i);
                __syncwarp();
            };
            mbar_wait(f_full, 0);
            UD_TR(leader, 17, 1);
            issue_scores(0);
            if (T2 > 1) issue_scores(1);
            for (int i = 0; i < T2; ++i) {
                const int s = i % NST, bf = i & 1;
                const uint32_t ph = (i >> 1) & 1;
                const uint32_t aSA = smem_u32(sST + s * 2 * S::SUB_BYTES), aSB = aSA + S::SUB_BYTES;
                const uint32_t tSc = tmem + bf * 128, tDp = tSc + 64;
                const uint32_t acc0 = i != 0;
                if (MODE == 0) {
                    // dV += P^T dO as soon as P^T is stored: it runs while the warpgroup still forms dS^T
                    mbar_wait(&p_rdy[bf], ph);
                    tc_fence_after();
                    UD_TR(leader, 2, i);
                    if (leader) {
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) umma_ts(tAcc1, tSc + ks * 8, desc_mnmajor64(aSB, ks), idesc_acc, acc0 | (ks != 0));
                    }
                    __syncwarp();
                }
                mbar_wait(&ds_rdy[bf], ph);
                tc_fence_after();
                UD_TR(leader, 3, i);
                if (leader) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) umma_ts(tAcc0, tDp + ks * 8, desc_mnmajor64(aSA, ks), idesc_acc, acc0 | (ks != 0));
                    umma_commit(&st_empty[s]);
                    umma_commit(&acc_done[bf]);
                    if (i + 1 == T2) umma_commit(all_done);
                }
                UD_TR(leader, 4, i);
                __syncwarp();
                if (i + 2 < T2) {
                    // scores(i+2) overwrite the columns the accumulation MMAs above read P / dS from.  tcgen05.mma issued by
                    // one thread execute in issue order, so no completion wait is needed here (UD_ATTN_BWD_SAFE=1 re-enables it)
                    if (p.safe_order) mbar_wait(&acc_done[bf], ph);
                    issue_scores(i + 2);
                }
            }
        }
    } else {
        // two softmax warpgroups (warps 2-5 and 6-9): warpgroup g owns the sub-tiles i == g (mod 2), i.e. score buffer g,
        // so two sub-tiles are in flight and every SM sub-partition has two warps to hide ALU / MUFU latency
        const int wg = (warp - 2) >> 2;
        const int qd = warp & 3;
        const int rloc = qd * 32 + lane;
        const int row = t0 + rloc;   // MODE0: key index ; MODE1: query index
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
        const int tid128 = (threadIdx.x - 64) & 127;
        const long long bh = (long long)b * p.H + h;
        int sid_row = 0;
        if (use_ids) sid_row = row < p.N ? (int)p.sample_ids[(long long)b * p.N + row] : -1;
        float lse_row = 0.f, dlt_row = 0.f;
        if (MODE == 1 && row < p.N) {
            lse_row = p.lse[bh * p.N + row] * LOG2E;
            if (p.o != nullptr) {
                // delta of this thread's query row, straight from O and dO (both warpgroups own the same rows and compute it
                // redundantly; warpgroup 0 publishes it)
                const uint4* po = reinterpret_cast<const uint4*>(p.o + ((long long)b * p.N + row) * p.ldo + h * HD);
                const uint4* pd = reinterpret_cast<const uint4*>(p.d_o + ((long long)b * p.N + row) * p.ldo + h * HD);
                float acc = 0.f;
#pragma unroll 4
                for (int j = 0; j < HD / 8; ++j) {
                    const uint4 a = po[j], g = pd[j];
                    acc += bf16lo(a.x) * bf16lo(g.x) + bf16hi(a.x) * bf16hi(g.x) + bf16lo(a.y) * bf16lo(g.y) + bf16hi(a.y) * bf16hi(g.y)
                         + bf16lo(a.z) * bf16lo(g.z) + bf16hi(a.z) * bf16hi(g.z) + bf16lo(a.w) * bf16lo(g.w) + bf16hi(a.w) * bf16hi(g.w);
                }
                dlt_row = acc;
                if (wg == 0) p.delta_out[bh * p.N + row] = acc;
            } else {
                dlt_row = p.delta[bh * p.N + row];
            }
        }
        const bool row_ok = row < p.N;
        const int bf = wg;
        const uint32_t tSc = tmem + bf * 128, tDp = tSc + 64;
        const float scl = p.scale_log2;
        const int Ntok = p.N;
        // metadata of this warpgroup's next sub-tile is fetched one iteration ahead (global latency off the critical path)
        float pf_lse = INFINITY, pf_dlt = 0.f;
        int pf_sid = -2;
        auto fetch_meta = [&](int ii) {
            if ((MODE == 0 || use_ids) && tid128 < 64 && ii < T2) {
                const int cidx = sub_of(ii) * 64 + tid128;
                if (MODE == 0) {
                    pf_lse = cidx < Ntok ? -p.lse[bh * Ntok + cidx] * LOG2E : -INFINITY;      // negated: added by the FFMA
                    pf_dlt = cidx < Ntok ? p.delta[bh * Ntok + cidx] : 0.f;
                }
                if (use_ids) pf_sid = cidx < Ntok ? (int)p.sample_ids[(long long)b * Ntok + cidx] : -2;
            }
        };
        fetch_meta(wg);
        for (int ii = wg; ii < T2; ii += 2) {
            const int i = sub_of(ii);                          // streamed sub-tile index (coordinates)
            const int ms = (wg * 2 + ((ii >> 1) & 1)) * 64;   // metadata slot of this sub-tile
            if (MODE == 0 || use_ids) {
                if (tid128 < 64) {
                    if (MODE == 0) { s_lse[ms + tid128] = pf_lse; s_dlt[ms + tid128] = pf_dlt; }
                    if (use_ids) s_sid[ms + tid128] = pf_sid;
                }
                named_bar_sync(1 + wg, 128);
                fetch_meta(ii + 2);
            }
            UD_TR(tid128 == 0, 14, ii);
            mbar_wait(&s_full[bf], (ii >> 1) & 1);
            tc_fence_after();
            UD_TR(tid128 == 0, 5, ii);
            // masks are needed only on edge tiles and on tile pairs that straddle a document boundary / hold padding
            const bool doc_mask = use_ids && !tl.nomask[ii];
            const bool slow = doc_mask || (i * 64 + 64 > Ntok) || !row_ok;
            uint32_t rsA[32], rsB[32], pk[32];
            tmem_ld_32x32b_x32(tSc + lane_off, rsA);
            tmem_ld_32x32b_x32(tSc + 32 + lane_off, rsB);
            tmem_ld_wait();
            UD_TR(tid128 == 0, 6, ii);
            // phase 1: P = exp(S * scale - lse), kept in fp32 in rsA / rsB (in place) and packed to bf16 in pk.  Interior tiles
            // take the mask-free instantiation (no per-element compare / select instructions).
            auto phase1 = [&](uint32_t (&rs)[32], int c, auto slow_tag) {
                constexpr bool SLOW = decltype(slow_tag)::value;
                const uint64_t scl2 = f2pack(scl, scl), nl_row2 = f2pack(-lse_row, -lse_row);
#pragma unroll
                for (int e4 = 0; e4 < 8; ++e4) {
                    uint64_t nl2[2] = {nl_row2, nl_row2};                      // -lse of the two column pairs
                    if (MODE == 0) {
                        const float4 lv = *reinterpret_cast<const float4*>(&s_lse[ms + c * 32 + e4 * 4]);
                        nl2[0] = f2pack(lv.x, lv.y); nl2[1] = f2pack(lv.z, lv.w);
                    }
                    float pv[4];
#pragma unroll
                    for (int h2 = 0; h2 < 2; ++h2) {
                        float x0, x1;
                        f2unpack(ffma2(f2pack(__uint_as_float(rs[e4 * 4 + 2 * h2]), __uint_as_float(rs[e4 * 4 + 2 * h2 + 1])), scl2, nl2[h2]), x0, x1);
                        pv[2 * h2] = ex2(x0); pv[2 * h2 + 1] = ex2(x1);
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (SLOW) {
                            const int col = c * 32 + e4 * 4 + u;
                            bool ok = row_ok && (i * 64 + col < Ntok);
                            if (doc_mask) {
                                const int sq = s_sid[ms + col];
                                ok = ok && sq == sid_row && sid_row != -1;
                            }
                            if (!ok) pv[u] = 0.f;
                        }
                        rs[e4 * 4 + u] = __float_as_uint(pv[u]);
                    }
                    if (MODE == 0) {
                        pk[c * 16 + e4 * 2] = pack_bf16x2(pv[0], pv[1]);
                        pk[c * 16 + e4 * 2 + 1] = pack_bf16x2(pv[2], pv[3]);
                    }
                }
            };
            if (slow) {
                phase1(rsA, 0, std::true_type{});
                phase1(rsB, 1, std::true_type{});
            } else {
                phase1(rsA, 0, std::false_type{});
                phase1(rsB, 1, std::false_type{});
            }
            UD_TR(tid128 == 0, 7, ii);
            if (MODE == 0) {
                tmem_st_32x32b_x32(tSc + lane_off, pk);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&p_rdy[bf]);
                UD_TR(tid128 == 0, 8, ii);
            }
            // phase 2: dS = P * (dP - delta)
            mbar_wait(&dp_full[bf], (ii >> 1) & 1);
            tc_fence_after();
            UD_TR(tid128 == 0, 9, ii);
            uint32_t rdA[32], rdB[32];
            tmem_ld_32x32b_x32(tDp + lane_off, rdA);
            tmem_ld_32x32b_x32(tDp + 32 + lane_off, rdB);
            tmem_ld_wait();
            UD_TR(tid128 == 0, 10, ii);
            auto phase2 = [&](const uint32_t (&rs)[32], const uint32_t (&rd)[32], int c) {
                const uint64_t d_row2 = f2pack(dlt_row, dlt_row);
#pragma unroll
                for (int e4 = 0; e4 < 8; ++e4) {
                    uint64_t d2[2] = {d_row2, d_row2};
                    if (MODE == 0) {
                        const float4 dv = *reinterpret_cast<const float4*>(&s_dlt[ms + c * 32 + e4 * 4]);
                        d2[0] = f2pack(dv.x, dv.y); d2[1] = f2pack(dv.z, dv.w);
                    }
#pragma unroll
                    for (int h2 = 0; h2 < 2; ++h2) {
                        const int e = e4 * 4 + 2 * h2;
                        float v0, v1;
                        f2unpack(fmul2(f2pack(__uint_as_float(rs[e]), __uint_as_float(rs[e + 1])),
                                       fsub2(f2pack(__uint_as_float(rd[e]), __uint_as_float(rd[e + 1])), d2[h2])), v0, v1);
                        pk[c * 16 + e4 * 2 + h2] = pack_bf16x2(v0, v1);
                    }
                }
            };
            phase2(rsA, rdA, 0);
            phase2(rsB, rdB, 1);
            UD_TR(tid128 == 0, 11, ii);
            tmem_st_32x32b_x32(tDp + lane_off, pk);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(&ds_rdy[bf]);
            UD_TR(tid128 == 0, 12, ii);
        }
        // ---- write the accumulators (the two warpgroups split the work) ----
        if (T2 > 0) mbar_wait(all_done, 0);
        tc_fence_after();
        UD_TR(tid128 == 0, 15, wg);
        {
            // MODE0: warpgroup 0 stores dK (Acc0), warpgroup 1 stores dV (Acc1).  MODE1: each stores half of dQ's columns.
            // TMEM gives every thread one ROW; the bf16 rows are staged in the (now idle) fixed-operand tiles with the
            // 16-byte chunks XOR-swizzled by row, then copied out with whole rows per warp instruction (coalesced 128-byte
            // lines instead of 32 scattered 16-byte stores per instruction).
            const int a = (MODE == 0) ? wg : 0;
            const uint32_t tA = a == 0 ? tAcc0 : tAcc1;
            uint8_t* stg = (MODE == 0 && wg == 1) ? sFB : sFA;
            const float sc = a == 0 ? p.scale : 1.0f;
            constexpr int CPR = HD * 2 / 16;                         // 16-byte chunks per row
            const int c_lo = (MODE == 0) ? 0 : wg * (HD / 64), c_hi = (MODE == 0) ? HD / 32 : (wg + 1) * (HD / 64);
#pragma unroll 1
            for (int c = c_lo; c < c_hi; ++c) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tA + c * 32 + lane_off, r);
                tmem_ld_wait();
                if (T2 == 0) {                                  // no matching tile at all (padding): the gradient is zero
#pragma unroll
                    for (int e = 0; e < 32; ++e) r[e] = 0u;
                }
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    uint4 o4;
                    o4.x = pack_bf16x2(__uint_as_float(r[8 * q4 + 0]) * sc, __uint_as_float(r[8 * q4 + 1]) * sc);
                    o4.y = pack_bf16x2(__uint_as_float(r[8 * q4 + 2]) * sc, __uint_as_float(r[8 * q4 + 3]) * sc);
                    o4.z = pack_bf16x2(__uint_as_float(r[8 * q4 + 4]) * sc, __uint_as_float(r[8 * q4 + 5]) * sc);
                    o4.w = pack_bf16x2(__uint_as_float(r[8 * q4 + 6]) * sc, __uint_as_float(r[8 * q4 + 7]) * sc);
                    const int ch = c * 4 + q4;
                    *reinterpret_cast<uint4*>(stg + rloc * (HD * 2) + ((ch ^ (rloc & (CPR - 1))) << 4)) = o4;
                }
            }
            named_bar_sync(1 + wg, 128);
            const int chunks_w = (c_hi - c_lo) * 4, chunk0 = c_lo * 4;            // this warpgroup's 16-byte chunks per row
            const long long ldo = a == 0 ? p.ld0 : p.ld1;
            __nv_bfloat16* base = (a == 0 ? p.out0 : p.out1) + ((long long)b * p.N + t0) * ldo + h * HD;
#pragma unroll 4
            for (int it = 0; it < chunks_w; ++it) {
                const int idx = it * 128 + tid128;
                const int rr = idx / chunks_w, ch = chunk0 + idx % chunks_w;
                if (t0 + rr < p.N) {
                    const uint4 v = *reinterpret_cast<const uint4*>(stg + rr * (HD * 2) + ((ch ^ (rr & (CPR - 1))) << 4));
                    *reinterpret_cast<uint4*>(base + (long long)rr * ldo + ch * 8) = v;
                }
            }
            UD_TR(tid128 == 0, 16, wg);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem);
    }
}

// ------------------------------------------------------------------------------------------------
// Backward v3.  Measured on the v2 kernels with in-kernel clock stamps (tools/attn_trace.py) and tools/ubench: a 128xNx16
// tcgen05.mma costs ~47 cycles at N=64 against ~65 at N=128 whether its A operand comes from shared memory or TMEM, so the
// 64-row streamed sub-tiles of v2 pay 72 % of the tensor time for 50 % of the work, and every sub-tile costs four
// warpgroup <-> tensor-pipe hand-offs.  v3 streams full 128-row tiles (all MMAs are 128x128x16) and lets BOTH softmax
// warpgroups work on every tile, each on 64 of its 128 columns:
//   dK/dV kernel (CTA owns key tile j, streams query tiles i), TMEM = St/Pt[128] | dPt/dSt[128] | dK[HD] | dV[HD]
//       issue order  dV(i), St(i+1), dK(i), dPt(i+1):  St(i+1) runs under the dS arithmetic of tile i
//   dQ kernel (CTA owns query tile i, streams key tiles j),   TMEM = S0[128] | S1[128] | dP[128] | dQ[HD]
//       S is double buffered (dS is written over the S it came from), dP single buffered but released as soon as the
//       warpgroups hold it in registers:  issue order  dP(j+1), dQ(j), S(j+2)
// The streamed operand that is read twice a tile apart (Q_i: St(i) ... dK(i);  K_j: S(j) ... dQ(j)) sits in a 3-deep ring,
// the other one (dO_i / V_j) in a 2-deep ring: 2 fixed + 5 streamed 32 KB tiles = the whole 227 KB.  Per-column metadata
// (log-sum-exp in the exp2 domain, delta, sample ids) is read straight from global memory with warp-uniform 16-byte loads
// (L1 hits): no shared staging, no named barriers in the loop.
// ------------------------------------------------------------------------------------------------
template <int HD>
struct AttnBwd3Smem {
    static constexpr int TILE = 128 * HD * 2;
    static constexpr int NA = 3, NB = 2;
    static constexpr int NBOX = HD / 64;
    static constexpr int NTILE = 2 + NA + NB;
    static constexpr int META = 2048;                    // barriers, TMEM pointer, document tile list
    static constexpr int BYTES = NTILE * TILE + 1024 + META;
};
static constexpr int MAX_DOC_TILES3 = 256;
using DocTiles3 = DocTilesT<MAX_DOC_TILES3>;

// TMEM column (relative to the tile's first column) of the bf16 A operand of reduction step ks: each warpgroup packs its CW
// probabilities into the first CW/2 columns of its own CW-column slice of the tile
template <int CW>   // CW = streamed columns per warpgroup (64 with two warpgroups, 32 with four)
UD_DEVINL uint32_t bwd3_pcol(int ks) { return (uint32_t)((ks / (CW / 16)) * CW + (ks % (CW / 16)) * 8); }

// one elected lane per warp signals on behalf of its 32 rows (after every lane's TMEM stores / loads have completed)
UD_DEVINL void warp_arrive(uint64_t* bar, int lane) {
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}

// P = exp2(S * scl - lse) for the 32-column chunk c of this thread's columns (fp32 kept in rs, for the dS product); SLOW applies
// the edge / document masks.  COLMETA: lse varies per column (dK/dV kernel) and is read (negated, exp2 domain) from nlse_col.
template <bool COLMETA, bool SLOW, int NPK>
UD_DEVINL void bwd3_probs(uint32_t (&rs)[32], uint32_t (&pk)[NPK], int c, float scl, float lse_row,
                          const float* __restrict__ nlse_col, int col0, int Ntok, bool row_ok, bool doc_mask, int sid_row,
                          const int64_t* __restrict__ sid_col) {
    const uint64_t scl2 = f2pack(scl, scl), nl_row2 = f2pack(-lse_row, -lse_row);
#pragma unroll
    for (int e4 = 0; e4 < 8; ++e4) {
        uint64_t nl2[2] = {nl_row2, nl_row2};                  // -lse of the two column pairs
        if (COLMETA) {
            if (!SLOW) {
                const float4 lv = __ldg(reinterpret_cast<const float4*>(nlse_col + c * 32 + e4 * 4));
                nl2[0] = f2pack(lv.x, lv.y); nl2[1] = f2pack(lv.z, lv.w);
            } else {
                float l4[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) l4[u] = (col0 + c * 32 + e4 * 4 + u < Ntok) ? __ldg(nlse_col + c * 32 + e4 * 4 + u) : -INFINITY;
                nl2[0] = f2pack(l4[0], l4[1]); nl2[1] = f2pack(l4[2], l4[3]);
            }
        }
        float pv[4];
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
            float x0, x1;
            f2unpack(ffma2(f2pack(__uint_as_float(rs[e4 * 4 + 2 * h2]), __uint_as_float(rs[e4 * 4 + 2 * h2 + 1])), scl2, nl2[h2]), x0, x1);
            pv[2 * h2] = ex2(x0); pv[2 * h2 + 1] = ex2(x1);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (SLOW) {
                const int col = col0 + c * 32 + e4 * 4 + u;
                bool ok = row_ok && col < Ntok;
                if (doc_mask && ok) ok = (int)__ldg(sid_col + c * 32 + e4 * 4 + u) == sid_row && sid_row != -1;
                if (!ok) pv[u] = 0.f;
            }
            rs[e4 * 4 + u] = __float_as_uint(pv[u]);
        }
        if (COLMETA) {       // dK/dV kernel: the bf16 probabilities are an MMA operand themselves
            pk[c * 16 + e4 * 2] = pack_bf16x2(pv[0], pv[1]);
            pk[c * 16 + e4 * 2 + 1] = pack_bf16x2(pv[2], pv[3]);
        }
    }
}

// dS = P * (dP - delta) for the 32-column chunk c of this thread's 64 columns, packed to bf16 pairs
template <bool COLMETA, bool SLOW, int NPK>
UD_DEVINL void bwd3_ds(const uint32_t (&rs)[32], const uint32_t (&rd)[32], uint32_t (&pk)[NPK], int c, float dlt_row,
                       const float* __restrict__ dlt_col, int col0, int Ntok) {
    const uint64_t d_row2 = f2pack(dlt_row, dlt_row);
#pragma unroll
    for (int e4 = 0; e4 < 8; ++e4) {
        uint64_t d2[2] = {d_row2, d_row2};
        if (COLMETA) {
            if (!SLOW) {
                const float4 dv = __ldg(reinterpret_cast<const float4*>(dlt_col + c * 32 + e4 * 4));
                d2[0] = f2pack(dv.x, dv.y); d2[1] = f2pack(dv.z, dv.w);
            } else {
                float d4[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) d4[u] = (col0 + c * 32 + e4 * 4 + u < Ntok) ? __ldg(dlt_col + c * 32 + e4 * 4 + u) : 0.f;
                d2[0] = f2pack(d4[0], d4[1]); d2[1] = f2pack(d4[2], d4[3]);
            }
        }
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
            const int e = e4 * 4 + 2 * h2;
            float v0, v1;
            f2unpack(fmul2(f2pack(__uint_as_float(rs[e]), __uint_as_float(rs[e + 1])),
                           fsub2(f2pack(__uint_as_float(rd[e]), __uint_as_float(rd[e + 1])), d2[h2])), v0, v1);
            pk[c * 16 + e4 * 2 + h2] = pack_bf16x2(v0, v1);
        }
    }
}

// accumulator -> bf16 rows staged (XOR-swizzled 16-byte chunks) in an idle fixed-operand tile -> coalesced global stores.
// One warpgroup (128 threads, thread = row) writes columns [c_lo*32, c_hi*32) of a [128][HD] accumulator.
template <int HD>
UD_DEVINL void bwd3_store_acc(uint32_t tA, uint32_t lane_off, uint8_t* stg, int rloc, int tid128, int bar_id, int c_lo, int c_hi,
                              float sc, bool zero, __nv_bfloat16* base, long long ldo, int rows_left) {
    constexpr int CPR = HD * 2 / 16;
#pragma unroll 1
    for (int c = c_lo; c < c_hi; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tA + c * 32 + lane_off, r);
        tmem_ld_wait();
        if (zero) {
#pragma unroll
            for (int e = 0; e < 32; ++e) r[e] = 0u;
        }
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
            uint4 o4;
            o4.x = pack_bf16x2(__uint_as_float(r[8 * q4 + 0]) * sc, __uint_as_float(r[8 * q4 + 1]) * sc);
            o4.y = pack_bf16x2(__uint_as_float(r[8 * q4 + 2]) * sc, __uint_as_float(r[8 * q4 + 3]) * sc);
            o4.z = pack_bf16x2(__uint_as_float(r[8 * q4 + 4]) * sc, __uint_as_float(r[8 * q4 + 5]) * sc);
            o4.w = pack_bf16x2(__uint_as_float(r[8 * q4 + 6]) * sc, __uint_as_float(r[8 * q4 + 7]) * sc);
            const int ch = c * 4 + q4;
            *reinterpret_cast<uint4*>(stg + rloc * (HD * 2) + ((ch ^ (rloc & (CPR - 1))) << 4)) = o4;
        }
    }
    named_bar_sync(bar_id, 128);
    const int chunks_w = (c_hi - c_lo) * 4, chunk0 = c_lo * 4;
#pragma unroll 4
    for (int it = 0; it < chunks_w; ++it) {
        const int idx = it * 128 + tid128;
        const int rr = idx / chunks_w, ch = chunk0 + idx % chunks_w;
        if (rr < rows_left) {
            const uint4 v = *reinterpret_cast<const uint4*>(stg + rr * (HD * 2) + ((ch ^ (rr & (CPR - 1))) << 4));
            *reinterpret_cast<uint4*>(base + (long long)rr * ldo + ch * 8) = v;
        }
    }
}

// MODE 0: dK/dV (fixed K_j, V_j; ring A = Q_i, ring B = dO_i).  MODE 1: dQ (fixed Q_i, dO_i; ring A = K_j, ring B = V_j).
// NWG softmax warpgroups share every tile, 128 / NWG columns each.  (Measured for the dQ kernel, whose warpgroups are the
// bottleneck: FOUR warpgroups of 32 columns each -- half the arithmetic per hand-off, four warps per SM sub-partition, 96
// registers -- are 2 % slower than two: 415.6 vs 406.0 us for the whole backward.)
template <int MODE> struct Bwd3Cfg { static constexpr int NWG = 2; static constexpr int THREADS = 64 + 128 * NWG; };
template <int HD, int MODE>
__global__ void __launch_bounds__(Bwd3Cfg<MODE>::THREADS, 1)
attn_bwd3_kernel(const __grid_constant__ CUtensorMap tm_fa, const __grid_constant__ CUtensorMap tm_fb,
                 const __grid_constant__ CUtensorMap tm_ra, const __grid_constant__ CUtensorMap tm_rb, const AttnBwdParams p) {
    using S = AttnBwd3Smem<HD>;
    constexpr int NA = S::NA, NB = S::NB, TILE = S::TILE;
    constexpr int NWG = Bwd3Cfg<MODE>::NWG, CW = 128 / NWG, NCH = CW / 32;      // columns / 32-column chunks per warpgroup
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sFA = smem;
    uint8_t* sFB = sFA + TILE;
    uint8_t* sRA = sFB + TILE;                  // [NA]
    uint8_t* sRB = sRA + NA * TILE;             // [NB]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sRB + NB * TILE);
    uint64_t* f_full = bars;                   // 1
    uint64_t* a_full = bars + 1;               // NA
    uint64_t* a_empty = a_full + NA;           // NA
    uint64_t* b_full = a_empty + NA;           // NB
    uint64_t* b_empty = b_full + NB;           // NB
    uint64_t* s_full = b_empty + NB;           // 2 (MODE 0 uses [0])   scores in TMEM
    uint64_t* dp_full = s_full + 2;            // 1
    uint64_t* p_rdy = dp_full + 1;             // 1   MODE 0: bf16 P^T stored          MODE 1: dP copied to registers (buffer free)
    uint64_t* ds_rdy = p_rdy + 1;              // 1   bf16 dS stored
    uint64_t* all_done = ds_rdy + 1;           // 1   completes exactly once (see v2)
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(all_done + 1);
    DocTiles3& tl = *reinterpret_cast<DocTiles3*>(tmem_ptr_smem + 4);
    static_assert((1 + 2 * NA + 2 * NB + 6) * 8 + 16 + sizeof(DocTiles3) <= S::META, "metadata area too small");

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;
    const int Tall = (p.N + 127) / 128;
    const bool use_ids = p.sample_ids != nullptr;
    UD_TR_INIT;
    UD_TR(threadIdx.x == 0, 17, 0);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_fa); tma_prefetch_desc(&tm_fb); tma_prefetch_desc(&tm_ra); tma_prefetch_desc(&tm_rb);
        mbar_init(f_full, 1);
        for (int s = 0; s < NA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < NB; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        mbar_init(&s_full[0], 1); mbar_init(&s_full[1], 1); mbar_init(dp_full, 1);
        mbar_init(p_rdy, 4 * NWG); mbar_init(ds_rdy, 4 * NWG);
        mbar_init(all_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_ptr_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr_smem;
    // MODE 0: R0 = St/Pt, R1 = dPt/dSt, accumulators dK | dV.   MODE 1: S0, S1, dP, accumulator dQ.
    const uint32_t tR0 = tmem, tR1 = tmem + 128;
    const uint32_t tAcc0 = MODE == 0 ? tmem + 256 : tmem + 384;       // dK | dQ
    const uint32_t tAcc1 = tmem + 256 + HD;                            // dV (MODE 0)
    const uint32_t tDP = MODE == 0 ? tR1 : tmem + 256;
    if (use_ids) doc_tile_list<128>(p.sample_ids + (long long)b * p.N, p.N, t0, Tall, tl);
    const int T = use_ids ? tl.n : Tall;
    auto tile_of = [&](int ii) { return use_ids ? (int)tl.idx[ii] : ii; };

    if (warp == 0) {
        if (lane == 0 && T > 0) {
            mbar_expect_tx(f_full, 2 * TILE);
#pragma unroll
            for (int bx = 0; bx < S::NBOX; ++bx) {
                tma_load_3d(sFA + bx * (128 * 128), &tm_fa, f_full, h * HD + bx * 64, t0, b);
                tma_load_3d(sFB + bx * (128 * 128), &tm_fb, f_full, h * HD + bx * 64, t0, b);
            }
            for (int i = 0; i < T; ++i) {
                const int sa = i % NA, sb = i % NB, t = tile_of(i);
                mbar_wait(&a_empty[sa], ((i / NA) & 1) ^ 1);
                UD_TR(true, 13, i);
                mbar_expect_tx(&a_full[sa], TILE);
#pragma unroll
                for (int bx = 0; bx < S::NBOX; ++bx)
                    tma_load_3d(sRA + sa * TILE + bx * (128 * 128), &tm_ra, &a_full[sa], h * HD + bx * 64, t * 128, b);
                mbar_wait(&b_empty[sb], ((i / NB) & 1) ^ 1);
                mbar_expect_tx(&b_full[sb], TILE);
#pragma unroll
                for (int bx = 0; bx < S::NBOX; ++bx)
                    tma_load_3d(sRB + sb * TILE + bx * (128 * 128), &tm_rb, &b_full[sb], h * HD + bx * 64, t * 128, b);
            }
        }
    } else if (warp == 1) {
        if (T > 0) {
            const uint32_t leader = elect_one();
            constexpr uint32_t idesc_sc = make_idesc_bf16(128, 128, false, false);
            constexpr uint32_t idesc_acc = make_idesc_bf16(128, HD, false, true);
            const uint32_t aFA = smem_u32(sFA), aFB = smem_u32(sFB), aRA = smem_u32(sRA), aRB = smem_u32(sRB);
            // scores of tile i from ring A (MODE 0: St = K_j Q_i^T -> R0;  MODE 1: S = Q_i K_j^T -> S[i & 1])
            auto issue_s = [&](int i) {
                const int sa = i % NA;
                mbar_wait(&a_full[sa], (i / NA) & 1);
                tc_fence_after();
                UD_TR(leader, 0, i);
                const uint32_t tS = MODE == 0 ? tR0 : tmem + (i & 1) * 128;
                if (leader) {
#pragma unroll
                    for (int ks = 0; ks < HD / 16; ++ks) umma_ss(tS, desc_kmajor(aFA, ks), desc_kmajor(aRA + sa * TILE, ks), idesc_sc, ks != 0);
                    umma_commit(&s_full[MODE == 0 ? 0 : (i & 1)]);
                }
                UD_TR(leader, 1, i);
                __syncwarp();
            };
            // dP of tile i from ring B (MODE 0: dPt = V_j dO_i^T;  MODE 1: dP = dO_i V_j^T).  MODE 1: the last read of V_j.
            auto issue_dp = [&](int i) {
                const int sb = i % NB;
                mbar_wait(&b_full[sb], (i / NB) & 1);
                tc_fence_after();
                UD_TR(leader, 18, i);
                if (leader) {
#pragma unroll
                    for (int ks = 0; ks < HD / 16; ++ks) umma_ss(tDP, desc_kmajor(aFB, ks), desc_kmajor(aRB + sb * TILE, ks), idesc_sc, ks != 0);
                    umma_commit(dp_full);
                    if (MODE == 1) umma_commit(&b_empty[sb]);
                }
                UD_TR(leader, 19, i);
                __syncwarp();
            };
            mbar_wait(f_full, 0);
            UD_TR(leader, 17, 1);
            issue_s(0);
            issue_dp(0);
            if (MODE == 1 && T > 1) issue_s(1);
            for (int i = 0; i < T; ++i) {
                const int sa = i % NA, sb = i % NB;
                const uint32_t ph = i & 1;
                const uint32_t acc0 = i != 0;
                if (MODE == 0) {
                    mbar_wait(p_rdy, ph);
                    tc_fence_after();
                    UD_TR(leader, 2, i);
                    if (leader) {
                        // dV += Pt dO_i  (reduction over the 128 queries of the tile; dO_i re-read MN-major).  Last read of dO_i.
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks) umma_ts(tAcc1, tR0 + bwd3_pcol<CW>(ks), desc_mnmajor(aRB + sb * TILE, ks), idesc_acc, acc0 | (ks != 0));
                        umma_commit(&b_empty[sb]);
                    }
                    __syncwarp();
                    // St(i+1) overwrites Pt(i): tcgen05.mma issued by one thread execute in issue order (as v2 relies on)
                    if (i + 1 < T) issue_s(i + 1);
                } else {
                    if (i + 1 < T) {
                        mbar_wait(p_rdy, ph);            // both warpgroups hold dP(i) in registers
                        tc_fence_after();
                        UD_TR(leader, 2, i);
                        issue_dp(i + 1);
                    }
                }
                mbar_wait(ds_rdy, ph);
                tc_fence_after();
                UD_TR(leader, 3, i);
                if (leader) {
                    // MODE 0: dK += dSt Q_i.   MODE 1: dQ += dS K_j.   Last read of the ring-A tile.
                    const uint32_t tA = MODE == 0 ? tR1 : tmem + (i & 1) * 128;
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) umma_ts(tAcc0, tA + bwd3_pcol<CW>(ks), desc_mnmajor(aRA + sa * TILE, ks), idesc_acc, acc0 | (ks != 0));
                    umma_commit(&a_empty[sa]);
                    if (i + 1 == T) umma_commit(all_done);
                }
                UD_TR(leader, 4, i);
                __syncwarp();
                if (MODE == 0) {
                    if (i + 1 < T) issue_dp(i + 1);      // overwrites dSt(i) after dK(i) has read it (issue order)
                } else {
                    if (i + 2 < T) issue_s(i + 2);       // overwrites dS(i) after dQ(i) has read it
                }
            }
        }
    } else {
        // NWG warpgroups of four warps; thread = accumulator row (MODE 0: key, MODE 1: query); warpgroup w handles the CW streamed
        // columns [CW*w, CW*w + CW) of every tile, as NCH chunks of 32
        const int wg = (warp - 2) >> 2;
        const int qd = warp & 3;
        const int rloc = qd * 32 + lane;
        const int row = t0 + rloc;
        const uint32_t lane_off = (uint32_t)(qd * 32) << 16;
        const int tid128 = (threadIdx.x - 64) & 127;
        const long long bh = (long long)b * p.H + h;
        const int Ntok = p.N;
        const bool row_ok = row < Ntok;
        int sid_row = 0;
        if (use_ids) sid_row = row_ok ? (int)p.sample_ids[(long long)b * Ntok + row] : -1;
        float lse_row = 0.f, dlt_row = 0.f;
        if (MODE == 1 && row_ok) {
            lse_row = p.lse[bh * Ntok + row] * LOG2E;
            dlt_row = p.delta[bh * Ntok + row];
        }
        const float scl = p.scale_log2;
        const uint32_t cb = (uint32_t)(wg * CW);
        const bool vec_ok = (Ntok & 3) == 0;           // 16-byte metadata loads need aligned per-head rows
        // MODE 0: the per-column metadata of a tile (CW lse2 + CW delta floats per warpgroup, 128-byte lines) is pulled into L1 one
        // tile ahead by the first lanes of every warp, so the warp-uniform loads in the softmax loops hit L1 (~35 cycles) instead
        // of paying an L2 / DRAM round trip per tile on the critical path
        auto prefetch_meta = [&](int ii) {
            if (MODE == 0 && ii < T && lane < 2 * NCH) {
                const int c0 = tile_of(ii) * 128 + wg * CW + (lane % NCH) * 32;
                if (c0 < Ntok) {
                    const float* src = ((lane >= NCH) ? p.delta : p.lse2) + bh * Ntok + c0;
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(src));
                }
            }
        };
        prefetch_meta(0);
        uint32_t rs[NCH][32], pk[NCH * 16];
        auto store_pk = [&](uint32_t taddr) {
            if constexpr (NCH == 2) tmem_st_32x32b_x32(taddr, pk); else tmem_st_32x32b_x16(taddr, pk);
        };
        for (int ii = 0; ii < T; ++ii) {
            prefetch_meta(ii + 1);
            const int t = tile_of(ii);
            const int col0 = t * 128 + wg * CW;       // first streamed index (MODE 0: query, MODE 1: key) of this warpgroup
            const bool doc_mask = use_ids && !tl.nomask[ii];
            const bool slow = doc_mask || (t * 128 + 128 > Ntok) || !row_ok || (MODE == 0 && !vec_ok);
            const float* lse_col = MODE == 0 ? p.lse2 + bh * Ntok + col0 : nullptr;
            const float* dlt_col = MODE == 0 ? p.delta + bh * Ntok + col0 : nullptr;
            const int64_t* sid_col = use_ids ? p.sample_ids + (long long)b * Ntok + col0 : nullptr;
            const uint32_t tS = (MODE == 0 ? tR0 : tmem + (ii & 1) * 128) + cb;
            UD_TR(tid128 == 0, 14, ii + 16 * wg);
            // (requesting S(ii+1) at the end of tile ii, to hide the TMEM read latency, was measured: it only moves the ~340 cycles
            //  from the head of a tile to its tail, where they delay the dS hand-off instead)
            mbar_wait(&s_full[MODE == 0 ? 0 : (ii & 1)], MODE == 0 ? (ii & 1) : ((ii >> 1) & 1));
            tc_fence_after();
            UD_TR(tid128 == 0, 5, ii + 16 * wg);
#pragma unroll
            for (int c = 0; c < NCH; ++c) tmem_ld_32x32b_x32(tS + c * 32 + lane_off, rs[c]);
            tmem_ld_wait();
            UD_TR(tid128 == 0, 6, ii + 16 * wg);
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                if (slow) bwd3_probs<MODE == 0, true>(rs[c], pk, c, scl, lse_row, lse_col, col0, Ntok, row_ok, doc_mask, sid_row, sid_col);
                else bwd3_probs<MODE == 0, false>(rs[c], pk, c, scl, lse_row, lse_col, col0, Ntok, row_ok, doc_mask, sid_row, sid_col);
            }
            UD_TR(tid128 == 0, 7, ii + 16 * wg);
            if (MODE == 0) {
                store_pk(tS + lane_off);
                tmem_st_wait();
                warp_arrive(p_rdy, lane);
                UD_TR(tid128 == 0, 8, ii + 16 * wg);
            }
            mbar_wait(dp_full, ii & 1);
            tc_fence_after();
            UD_TR(tid128 == 0, 9, ii + 16 * wg);
            if (MODE == 0) {
                // one 32-column chunk of dPt at a time (the fp32 probabilities stay live: register budget)
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    uint32_t rd[32];
                    tmem_ld_32x32b_x32(tDP + cb + c * 32 + lane_off, rd);
                    tmem_ld_wait();
                    if (slow) bwd3_ds<true, true>(rs[c], rd, pk, c, dlt_row, dlt_col, col0, Ntok);
                    else bwd3_ds<true, false>(rs[c], rd, pk, c, dlt_row, dlt_col, col0, Ntok);
                }
                UD_TR(tid128 == 0, 11, ii + 16 * wg);
            } else {
                uint32_t rd[NCH][32];
#pragma unroll
                for (int c = 0; c < NCH; ++c) tmem_ld_32x32b_x32(tDP + cb + c * 32 + lane_off, rd[c]);
                tmem_ld_wait();
                warp_arrive(p_rdy, lane);              // dP buffer may be overwritten by dP(j+1)
                UD_TR(tid128 == 0, 10, ii + 16 * wg);
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    if (slow) bwd3_ds<false, true>(rs[c], rd[c], pk, c, dlt_row, dlt_col, col0, Ntok);
                    else bwd3_ds<false, false>(rs[c], rd[c], pk, c, dlt_row, dlt_col, col0, Ntok);
                }
                UD_TR(tid128 == 0, 11, ii + 16 * wg);
            }
            // MODE 0: dSt over dPt (R1).  MODE 1: dS over the S buffer it came from.
            store_pk((MODE == 0 ? tR1 + cb : tS) + lane_off);
            tmem_st_wait();
            warp_arrive(ds_rdy, lane);
            UD_TR(tid128 == 0, 12, ii + 16 * wg);
        }
        if (T > 0) mbar_wait(all_done, 0);
        tc_fence_after();
        UD_TR(tid128 == 0, 15, wg);
        const int rows_left = Ntok - t0;
        if (MODE == 0) {
            // warpgroup 0 stores dK (scaled), warpgroup 1 stores dV
            const uint32_t tA = wg == 0 ? tAcc0 : tAcc1;
            __nv_bfloat16* base = (wg == 0 ? p.out0 : p.out1) + ((long long)b * Ntok + t0) * (wg == 0 ? p.ld0 : p.ld1) + h * HD;
            bwd3_store_acc<HD>(tA, lane_off, wg == 0 ? sFA : sFB, rloc, tid128, 1 + wg, 0, HD / 32, wg == 0 ? p.scale : 1.0f, T == 0,
                               base, wg == 0 ? p.ld0 : p.ld1, rows_left);
        } else {
            // the 32-column chunks of dQ are dealt round-robin to the warpgroups (disjoint chunks of one staging tile)
            constexpr int NC = HD / 32, PER = (NC + NWG - 1) / NWG;
            const int c_lo = min(wg * PER, NC), c_hi = min(c_lo + PER, NC);
            __nv_bfloat16* base = p.out0 + ((long long)b * Ntok + t0) * p.ld0 + h * HD;
            bwd3_store_acc<HD>(tAcc0, lane_off, sFA, rloc, tid128, 1 + wg, c_lo, c_hi, p.scale, T == 0, base, p.ld0, rows_left);
        }
        UD_TR(tid128 == 0, 16, wg);
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

static int make_head_tmap(CUtensorMap* tm, const void* base, long long ld, int B, int N, int D, int box_rows = 128,
                          long long batch_stride = 0) {
    // dims: (column, token, batch); box = 64 columns x box_rows tokens x 1 batch.  batch_stride (elements) defaults to N * ld.
    return make_tmap_3d_bf16(tm, base, (uint64_t)D, (uint64_t)N, (uint64_t)B, (uint64_t)ld,
                             (uint64_t)(batch_stride > 0 ? batch_stride : ld * N), 64, box_rows, 1);
}

template <int HD>
static int launch_attn_fwd(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv,
                           const AttnParams& p, cudaStream_t stream, long long q_bs = 0, long long k_bs = 0, long long v_bs = 0) {
    CUtensorMap tq, tk, tv;
    const int D = p.H * HD;
    int rc = make_head_tmap(&tq, q, ldq, p.B, p.N, D, 128, q_bs);
    if (rc) return rc;
    if ((rc = make_head_tmap(&tk, k, ldk, p.B, p.Nk, D, 128, k_bs))) return rc;
    if ((rc = make_head_tmap(&tv, v, ldv, p.B, p.Nk, D, 128, v_bs))) return rc;
    const int smem3 = attn_smem_bytes<HD>(5);
    static bool attr3 = false;
    const int smem5 = 3 * AttnSmem<HD>::TILE_BYTES + 1024 + 256;
    CUtensorMap tk64, tv64;
    if ((rc = make_head_tmap(&tk64, k, ldk, p.B, p.Nk, D, 64, k_bs))) return rc;
    if ((rc = make_head_tmap(&tv64, v, ldv, p.B, p.Nk, D, 64, v_bs))) return rc;
    if (!attr3) {
        UD_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd6_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem5));
        UD_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd3_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3));
        attr3 = true;
    }
    // A/B switch: 3 = one CTA/SM, double-buffered 128-key S (round 1); 6 = two CTAs/SM, double-buffered 64-key S (default)
    const char* fwd_env = getenv("UD_ATTN_FWD");           // read per call (see UD_ATTN_BWD)
    const int variant = fwd_env != nullptr ? atoi(fwd_env) : 6;
    dim3 grid3((p.N + ATT_BQ - 1) / ATT_BQ, p.H, p.B);
    if (variant == 3) attn_fwd3_kernel<HD><<<grid3, 320, smem3, stream>>>(tq, tk, tv, p);
    else attn_fwd6_kernel<HD><<<grid3, 192, smem5, stream>>>(tq, tk64, tv64, p);
    UD_CUDA_CHECK(cudaGetLastError());
    return 0;
}

template <int HD>
static int launch_attn_bwd(const void* q, const void* k, long long ldqk, const void* v, long long ldv, const void* d_o,
                           long long ldo, AttnBwdParams p, __nv_bfloat16* dq, __nv_bfloat16* dk, long long lddqk,
                           __nv_bfloat16* dv, long long lddv, cudaStream_t stream) {
    // Default: the v3 dQ kernel (128-row streamed tiles, double-buffered S) with the v2 dK/dV kernel (64-row sub-tiles, two
    // sub-tiles in flight) -- the faster of each pair at B8 H16 N1280 hd128.  UD_ATTN_BWD=2 / 3 force one generation for both.
    // Packed (document-masked) batches keep the v2 dQ kernel: its per-column sample ids come from shared memory, v3 reads them
    // from global memory on every masked tile (cfg5, same box: 89.2 ms per step vs 90.7 with v3 forced).
    // (read per call, not cached: tests/test_attention_gpu.py runs every generation kept in the library in one process)
    const char* gen_env = getenv("UD_ATTN_BWD");
    const int gen = gen_env != nullptr ? atoi(gen_env) : 0;
    const bool dq_v3 = gen == 3 || (gen != 2 && p.sample_ids == nullptr);
    const bool dkv_v3 = gen == 3;
    // The v2 dQ kernel can form delta itself (UD_ATTN_FUSED_DELTA=1), but measured at B8 H16 N1280 hd128 the per-thread row reads
    // delay every CTA's first sub-tile: 458.6 us fused vs 445.4 us with the separate 33 us pass, so the pass stays the default.
    const bool sep_delta = dq_v3 || getenv("UD_ATTN_FUSED_DELTA") == nullptr;
    if (sep_delta) {
        const long long warps = (long long)p.B * p.N * p.H;
        const int threads = 256;
        const long long blocks = (warps * 32 + threads - 1) / threads;
        p.lse2 = dkv_v3 ? p.delta_out + warps : nullptr;         // second half of the caller's scratch (v3 dK/dV kernel only)
        attn_delta_kernel<<<(unsigned)blocks, threads, 0, stream>>>(p.o, p.d_o, p.ldo, p.delta_out, p.B, p.N, p.H, HD, p.lse,
                                                                    const_cast<float*>(p.lse2));
        UD_CUDA_CHECK(cudaGetLastError());
        p.o = nullptr;                        // kernels read p.delta
    }
    CUtensorMap tq, tk, tv, tdo, tq64, tk64, tv64, tdo64;
    const int D = p.H * HD;
    int rc;
    if ((rc = make_head_tmap(&tq, q, ldqk, p.B, p.N, D))) return rc;
    if ((rc = make_head_tmap(&tk, k, ldqk, p.B, p.N, D))) return rc;
    if ((rc = make_head_tmap(&tv, v, ldv, p.B, p.N, D))) return rc;
    if ((rc = make_head_tmap(&tdo, d_o, ldo, p.B, p.N, D))) return rc;
    dim3 grid((p.N + 127) / 128, p.H, p.B);
    AttnBwdParams p0 = p;
    p0.out0 = dk; p0.ld0 = lddqk; p0.out1 = dv; p0.ld1 = lddv;
    AttnBwdParams p1 = p;
    p1.out0 = dq; p1.ld0 = lddqk; p1.out1 = nullptr; p1.ld1 = 0;
#ifdef UD_ATTN_TRACE
    p0.trace = p1.trace = g_attn_trace_host;
    const char* only = getenv("UD_ATTN_TRACE_ONLY");      // tools/attn_trace.py: one kernel stamps the table at a time
    const bool run_dq = only == nullptr || only[1] == 'q', run_dkv = only == nullptr || only[1] == 'k';
#else
    constexpr bool run_dq = true, run_dkv = true;
#endif
    const int smem3 = AttnBwd3Smem<HD>::BYTES, smem2 = AttnBwd2Smem<HD>::BYTES;
    static bool attr_set = false;
    if (!attr_set) {
        UD_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd3_kernel<HD, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3));
        UD_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd3_kernel<HD, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3));
        UD_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd2_kernel<HD, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
        UD_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd2_kernel<HD, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2));
        attr_set = true;
    }
    if (!dq_v3 || !dkv_v3) {
        if ((rc = make_head_tmap(&tq64, q, ldqk, p.B, p.N, D, 64))) return rc;
        if ((rc = make_head_tmap(&tk64, k, ldqk, p.B, p.N, D, 64))) return rc;
        if ((rc = make_head_tmap(&tv64, v, ldv, p.B, p.N, D, 64))) return rc;
        if ((rc = make_head_tmap(&tdo64, d_o, ldo, p.B, p.N, D, 64))) return rc;
    }
    // dQ first (the fused-delta variant of the v2 dQ kernel also produces delta, which the dK/dV kernel reads)
    if (run_dq) {
        if (dq_v3) attn_bwd3_kernel<HD, 1><<<grid, Bwd3Cfg<1>::THREADS, smem3, stream>>>(tq, tdo, tk, tv, p1);
        else attn_bwd2_kernel<HD, 1><<<grid, 320, smem2, stream>>>(tq, tdo, tk64, tv64, p1);
    }
    UD_CUDA_CHECK(cudaGetLastError());
    if (run_dkv) {
        if (dkv_v3) attn_bwd3_kernel<HD, 0><<<grid, Bwd3Cfg<0>::THREADS, smem3, stream>>>(tk, tv, tq, tdo, p0);
        else attn_bwd2_kernel<HD, 0><<<grid, 320, smem2, stream>>>(tk, tv, tq64, tdo64, p0);
    }
    UD_CUDA_CHECK(cudaGetLastError());
    return 0;
}

}  // namespace ud

using namespace ud;

static int attn_fwd_entry(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, void* o,
                          long long ldo, float* lse, const int64_t* sample_ids, int B, int Nq, int Nk, int H, int head_dim,
                          float scale, void* stream, long long q_bs = 0, long long k_bs = 0, long long v_bs = 0, long long o_bs = 0) {
    if (B <= 0 || Nq <= 0 || Nk <= 0) return 0;
    if (sample_ids != nullptr && (Nq != Nk || (Nk + 63) / 64 > MAX_DOC_TILES)) {
        fprintf(stderr, "unidisc_b200: document-masked attention needs Nq == Nk <= %d\n", MAX_DOC_TILES * 64);
        return -1;
    }
    AttnParams p;
    p.B = B; p.N = Nq; p.Nk = Nk; p.H = H;
    p.scale_log2 = scale * LOG2E;
    p.o = reinterpret_cast<__nv_bfloat16*>(o);
    p.ldo = ldo;
    p.o_bs = o_bs > 0 ? o_bs : (long long)Nq * ldo;
    p.lse = lse;
    p.sample_ids = sample_ids;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (head_dim == 128) return launch_attn_fwd<128>(q, ldq, k, ldk, v, ldv, p, s, q_bs, k_bs, v_bs);
    if (head_dim == 64) return launch_attn_fwd<64>(q, ldq, k, ldk, v, ldv, p, s, q_bs, k_bs, v_bs);
    fprintf(stderr, "unidisc_b200: attention supports head_dim 64 and 128 (got %d)\n", head_dim);
    return -1;
}

extern "C" int ud_attn_fwd(const void* q, const void* k, long long ldqk, const void* v, long long ldv, void* o, long long ldo,
                           float* lse, const int64_t* sample_ids, int B, int N, int H, int head_dim, float scale, void* stream) {
    return attn_fwd_entry(q, ldqk, k, ldqk, v, ldv, o, ldo, lse, sample_ids, B, N, N, H, head_dim, scale, stream);
}

extern "C" int ud_attn_fwd_kv(const void* q, long long ldq, long long q_bs, const void* k, long long ldk, long long k_bs,
                              const void* v, long long ldv, long long v_bs, void* o, long long ldo, long long o_bs, float* lse, int B,
                              int Nq, int Nk, int H, int head_dim, float scale, void* stream) {
    return attn_fwd_entry(q, ldq, k, ldk, v, ldv, o, ldo, lse, nullptr, B, Nq, Nk, H, head_dim, scale, stream, q_bs, k_bs, v_bs, o_bs);
}

extern "C" int ud_attn_bwd(const void* q, const void* k, long long ldqk, const void* v, long long ldv, const void* o,
                           const void* d_o, long long ldo, const float* lse, float* delta, void* dq, void* dk, long long lddqk,
                           void* dv, long long lddv, const int64_t* sample_ids, int B, int N, int H, int head_dim, float scale,
                           void* stream) {
    if (B <= 0 || N <= 0) return 0;
    if (sample_ids != nullptr && (N + 127) / 128 > MAX_DOC_TILES3) {
        fprintf(stderr, "unidisc_b200: document-masked attention needs N <= %d\n", MAX_DOC_TILES3 * 128);
        return -1;
    }
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    AttnBwdParams p;
    p.o = reinterpret_cast<const __nv_bfloat16*>(o);
    p.d_o = reinterpret_cast<const __nv_bfloat16*>(d_o);
    p.ldo = ldo;
    p.delta_out = delta;
    p.B = B; p.N = N; p.H = H;
    p.scale = scale;
    p.scale_log2 = scale * LOG2E;
    p.lse = lse;
    p.delta = delta;
    p.sample_ids = sample_ids;
    p.out0 = p.out1 = nullptr;
    p.ld0 = p.ld1 = 0;
    p.safe_order = getenv("UD_ATTN_BWD_SAFE") != nullptr;
    p.lse2 = nullptr;
    auto* dqp = reinterpret_cast<__nv_bfloat16*>(dq);
    auto* dkp = reinterpret_cast<__nv_bfloat16*>(dk);
    auto* dvp = reinterpret_cast<__nv_bfloat16*>(dv);
    if (head_dim == 128) return launch_attn_bwd<128>(q, k, ldqk, v, ldv, d_o, ldo, p, dqp, dkp, lddqk, dvp, lddv, s);
    if (head_dim == 64) return launch_attn_bwd<64>(q, k, ldqk, v, ldv, d_o, ldo, p, dqp, dkp, lddqk, dvp, lddv, s);
    fprintf(stderr, "unidisc_b200: attention supports head_dim 64 and 128 (got %d)\n", head_dim);
    return -1;
}
