// tcgen05 GEMM family for the DiT block (sm_100a).
//
//   C[M,N] = sum_k A(m,k) * B(n,k)        bf16 operands, fp32 accumulation in TMEM
//
// Operand layouts (selected at run time, compiled as template variants):
//   ta=0  A given as [M,K] row-major ("K-major")        ta=1  A given as [K,M] row-major ("MN-major")
//   tb=0  B given as [N,K] row-major ("K-major")        tb=1  B given as [K,N] row-major ("MN-major")
// so that  forward  y = x W^T        is (ta=0,tb=0) with B = W [out,in]            (reference nn.Linear, dit.py:642,887,917-919)
//          dgrad    dx = dy W        is (ta=0,tb=1) with B = W [out,in] read as [K=out, N=in]
//          wgrad    dW = dy^T x      is (ta=1,tb=1) with A = dy [tokens,out], B = x [tokens,in]
// without any transposed copies of weights or activations.
//
// Structure: persistent CTAs (one per SM), 6 warps: warp0 = TMA producer, warp1 = UMMA issuer (+TMEM owner),
// warps 2..5 = epilogue (TMEM -> registers -> global).  128 x BN x 64 tiles, SWIZZLE_128B smem stages fed by TMA,
// two TMEM accumulator stages so the epilogue of tile i overlaps the main loop of tile i+1.
//
// Fused epilogues: +bias, GELU(tanh) (emits both pre-activation u and g), GELU-backward (multiplies by g'(u)),
// fp32 store / accumulate for weight gradients.
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "unidisc_b200.h"

namespace ud {

static constexpr int BM = 128;
static constexpr int BK = 64;
static constexpr int UMMA_K = 16;

// Tile order: groups of RASTER_GM row-tiles are swept across all column-tiles before moving on, so the operand panels a
// wave of CTAs touches (GM row panels + ~wave/GM column panels) stay L2-resident instead of streaming all of A per column.
static constexpr int RASTER_GM = 8;
UD_DEVINL void tile_coords(int tile, int num_m, int num_n, int& tm, int& tn) {
    const int per_group = RASTER_GM * num_n;
    const int g = tile / per_group, local = tile - g * per_group;
    const int m_base = g * RASTER_GM;
    const int gm = min(RASTER_GM, num_m - m_base);
    tm = m_base + local % gm;
    tn = local / gm;
}

struct GemmParams {
    int M, N, K;
    void* C;
    long long ldc;
    const __nv_bfloat16* bias;
    void* aux;
    long long ld_aux;
    int num_m_tiles, num_n_tiles;
    // stream-K tail (2-CTA kernel): tiles [0, sk_first_tile) are whole-tile work items dealt round-robin; the remaining
    // sk_tiles tiles' (tile, k-block) iterations are cut into equal contiguous ranges of sk_w k-blocks per cluster, so
    // every cluster finishes at the same time instead of idling through a partial last wave.
    int sk_first_tile, sk_tiles, sk_w;
    float* sk_ws;        // [clusters][2][BN][128] fp32 partial accumulators (one slot per cluster)
    int* sk_flags;       // [sk_tiles] arrivals (8 epilogue warps per contributing cluster), [sk_tiles] owner-done counters
};

// One work item of a cluster: a whole tile, or a k-range of a stream-K tile.
struct WorkItem {
    int tile, k0, k1;
    int kind;            // 0 whole tile, 1 owner of a split tile (has k-block 0, does the epilogue), 2 partial contributor
    int sk_t;            // stream-K tile index (kind != 0)
};
struct WorkIter {
    int next_tile, stride, first_sk, lo, hi, num_kb;
    UD_DEVINL WorkIter(const GemmParams& p, int cluster_id, int num_clusters, int num_kb_) {
        next_tile = cluster_id; stride = num_clusters; first_sk = p.sk_first_tile; num_kb = num_kb_;
        const long long total = (long long)p.sk_tiles * num_kb_;
        const long long l = (long long)cluster_id * p.sk_w;
        lo = (int)(l < total ? l : total);
        const long long h = l + p.sk_w;
        hi = (int)(h < total ? h : total);
    }
    UD_DEVINL bool next(WorkItem& w) {
        if (next_tile < first_sk) {
            w.tile = next_tile; w.k0 = 0; w.k1 = num_kb; w.kind = 0; w.sk_t = -1;
            next_tile += stride;
            return true;
        }
        if (lo < hi) {
            const int t = lo / num_kb, k0 = lo - t * num_kb;
            const int k1 = min(num_kb, k0 + (hi - lo));
            w.tile = first_sk + t; w.k0 = k0; w.k1 = k1; w.sk_t = t;
            w.kind = (k0 == 0) ? (k1 == num_kb ? 0 : 1) : 2;
            lo += k1 - k0;
            return true;
        }
        return false;
    }
};

// Warp-cooperative store of a [32 rows x 128 bytes] slab whose row `lane` is held by thread `lane` as 8 x 16-byte units.
// Row-per-thread stores touch 32 different 128-byte lines per instruction with 16 bytes each: ncu showed the L1->XBAR
// request path 73 % busy and twice the payload in sector traffic (each 16-byte store occupies a 32-byte sector slot).
// Here the slab is transposed through a swizzled 4 KB shared-memory buffer so that every store instruction writes 4 FULL
// lines (8 lanes x 16 B per row).  ACC: read-add-write (fp32 gradient accumulation), same coalesced pattern.
// rows_ok = number of valid rows in the slab (row0 + i < M), dst = address of (row0, col0), pitch in bytes.
template <bool ACC>
UD_DEVINL void store_slab_coalesced(uint8_t* buf, const uint4 (&u)[8], int lane, char* dst, long long pitch, int rows_ok) {
#pragma unroll
    for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4*>(buf + lane * 128 + ((j ^ (lane & 7)) << 4)) = u[j];
    __syncwarp();
    const int uu = lane & 7, r0 = lane >> 3;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int rr = i * 4 + r0;
        uint4 v = *reinterpret_cast<const uint4*>(buf + rr * 128 + ((uu ^ (rr & 7)) << 4));
        if (rr < rows_ok) {
            uint4* g = reinterpret_cast<uint4*>(dst + (long long)rr * pitch + uu * 16);
            if constexpr (ACC) {
                const uint4 o = *g;
                v.x = __float_as_uint(__uint_as_float(v.x) + __uint_as_float(o.x));
                v.y = __float_as_uint(__uint_as_float(v.y) + __uint_as_float(o.y));
                v.z = __float_as_uint(__uint_as_float(v.z) + __uint_as_float(o.z));
                v.w = __float_as_uint(__uint_as_float(v.w) + __uint_as_float(o.w));
            }
            *g = v;
        }
    }
    __syncwarp();
}

// Epilogue of the 2-CTA kernel for one 32-column chunk.  Differences from epilogue_chunk: every GemmParams field it needs is
// passed in registers (the per-chunk LDC of p.N / p.bias / p.C were 15 % of the kernel's stall samples), the bias comes from
// a per-tile shared-memory copy (fp32, loaded before the accumulator is ready) and the GELU-backward pre-activations are
// prefetched one chunk ahead by the caller (`ux`), so no global-load latency sits between tcgen05.ld and the stores.
struct EpiRegs {
    int N;
    long long ldc, ld_aux;
    char* C;
    char* aux;
    const __nv_bfloat16* bias;
    float alpha;
};
template <int EPI>
UD_DEVINL void epilogue_chunk2(const uint32_t (&r)[32], int row, bool row_ok, int col0, const EpiRegs& e, const float* bias_s,
                               const uint4* ux /* prefetched pre-activations of this chunk or nullptr */) {
    const bool full = (col0 + 32 <= e.N);
    if (!row_ok) return;
    if constexpr (EPI == UD_EPI_F32 || EPI == UD_EPI_F32_ACC) {
        float* cp = reinterpret_cast<float*>(e.C) + (long long)row * e.ldc + col0;
        if (full) {
            float4 o[8];
            if constexpr (EPI == UD_EPI_F32_ACC) {
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = *reinterpret_cast<float4*>(cp + 4 * j);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float4 v = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                                       __uint_as_float(r[4 * j + 3]));
                if constexpr (EPI == UD_EPI_F32_ACC) { v.x += o[j].x; v.y += o[j].y; v.z += o[j].z; v.w += o[j].w; }
                *reinterpret_cast<float4*>(cp + 4 * j) = v;
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if (col0 + j < e.N) {
                    float v = __uint_as_float(r[j]);
                    if constexpr (EPI == UD_EPI_F32_ACC) v += cp[j];
                    cp[j] = v;
                }
            }
        }
    } else {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (bias_s != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float4 b = *reinterpret_cast<const float4*>(bias_s + 4 * j);     // broadcast LDS.128 (zero beyond N)
                v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
            }
        }
        if constexpr (EPI == UD_EPI_BF16_SCALED) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = bf16_round(v[j]) * e.alpha;
        }
        __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(e.C) + (long long)row * e.ldc + col0;
        __nv_bfloat16* xp = reinterpret_cast<__nv_bfloat16*>(e.aux) + (long long)row * e.ld_aux + col0;
        if constexpr (EPI == UD_EPI_BF16_DGELU) {
            if (full) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint4 u = ux != nullptr ? ux[j] : ldg_stream(reinterpret_cast<const uint4*>(xp) + j);
                    v[8 * j + 0] *= gelu_tanh_grad(bf16lo(u.x)); v[8 * j + 1] *= gelu_tanh_grad(bf16hi(u.x));
                    v[8 * j + 2] *= gelu_tanh_grad(bf16lo(u.y)); v[8 * j + 3] *= gelu_tanh_grad(bf16hi(u.y));
                    v[8 * j + 4] *= gelu_tanh_grad(bf16lo(u.z)); v[8 * j + 5] *= gelu_tanh_grad(bf16hi(u.z));
                    v[8 * j + 6] *= gelu_tanh_grad(bf16lo(u.w)); v[8 * j + 7] *= gelu_tanh_grad(bf16hi(u.w));
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (col0 + j < e.N) v[j] *= gelu_tanh_grad(__bfloat162float(xp[j]));
            }
        }
        if (full) {
            uint32_t o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                *(reinterpret_cast<uint4*>(cp) + j) = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
            if constexpr (EPI == UD_EPI_BF16_GELU) {
                uint32_t g[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) g[j] = pack_bf16x2(gelu_tanh(bf16lo(o[j])), gelu_tanh(bf16hi(o[j])));
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    *(reinterpret_cast<uint4*>(xp) + j) = make_uint4(g[4 * j], g[4 * j + 1], g[4 * j + 2], g[4 * j + 3]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if (col0 + j < e.N) {
                    __nv_bfloat16 ub = __float2bfloat16_rn(v[j]);
                    cp[j] = ub;
                    if constexpr (EPI == UD_EPI_BF16_GELU) xp[j] = __float2bfloat16_rn(gelu_tanh(__bfloat162float(ub)));
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// 2-CTA variant (cta_group::2): a CTA pair on one TPC computes a 256 x BN tile.  Each CTA stages its own 128 rows of A
// and HALF of the B tile (BN/2 rows), the leader's single thread issues M=256 UMMAs that read both CTAs' shared memory,
// and each CTA's TMEM receives its 128 accumulator rows.  Per flop this halves the B-operand traffic from L2, which is
// what bounds the 1-CTA kernel (128x256: 0.0117 B/flop ~ 20 TB/s at peak vs ~12 TB/s of L2).
// ------------------------------------------------------------------------------------------------
// 2-CTA kernel: warp0 TMA producer, warp1 UMMA issuer, warps 2..9 epilogue.  EIGHT epilogue warps (two per TMEM lane quarter,
// each taking half of the tile's columns): the fused epilogues run ~30 instructions per element, which one warp per
// scheduler could not issue in the time the tensor cores need for a tile (ncu: GELU kernels were epilogue-bound).
static constexpr int GEMM2_EPI_WARPS = 8;
static constexpr int GEMM2_THREADS = 64 + 32 * GEMM2_EPI_WARPS;

template <int BN>
struct Gemm2Cfg {
    static constexpr int A_BYTES = BM * BK * 2;            // 128 rows of A per CTA
    static constexpr int B_BYTES = (BN / 2) * BK * 2;      // half of the B tile per CTA
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (BN == 256) ? 5 : 7;
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int BIAS_BYTES = 2 * BN * 4;          // fp32 bias of the current / next tile (epilogue staging)
    static constexpr int STORE_BYTES = GEMM2_EPI_WARPS * 4096;   // per-warp 32 rows x 128 B transposition buffer (coalesced stores)
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + BIAS_BYTES + STORE_BYTES;
};

template <bool A_MN, bool B_MN, int BN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM2_THREADS, 1)
gemm2_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const GemmParams p) {
    using Cfg = Gemm2Cfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* bias_smem = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES + 256);     // [2][BN]
    uint8_t* store_smem = smem + STAGES * Cfg::STAGE_BYTES + 256 + Cfg::BIAS_BYTES;           // [epi warps][4096]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;
    const int num_kb = (p.K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], 2 * GEMM2_EPI_WARPS);     // the epilogue warps of both CTAs (used in the leader only)
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm<Cfg::TMEM_COLS>(tmem_ptr_smem);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            WorkIter it(p, cluster_id, num_clusters, num_kb);
            WorkItem w;
            while (it.next(w)) {
                int tm, tn;
                tile_coords(w.tile, p.num_m_tiles, p.num_n_tiles, tm, tn);
                const int m0 = tm * 256 + (int)rank * BM;
                const int n0 = tn * BN + (int)rank * (BN / 2);
                for (int kb = w.k0; kb < w.k1; ++kb) {
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
                    uint8_t* sb = sa + Cfg::A_BYTES;
                    if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * Cfg::STAGE_BYTES);   // bytes of BOTH CTAs land on the leader's barrier
                    const int k0 = kb * BK;
                    if constexpr (!A_MN) {
                        tma_load_2d_2sm(sa, &tma_a, &full_bar[s], k0, m0);
                    } else {
#pragma unroll
                        for (int j = 0; j < BM / 64; ++j) tma_load_2d_2sm(sa + j * 8192, &tma_a, &full_bar[s], m0 + 64 * j, k0);
                    }
                    if constexpr (!B_MN) {
                        tma_load_2d_2sm(sb, &tma_b, &full_bar[s], k0, n0);
                    } else {
#pragma unroll
                        for (int j = 0; j < BN / 128; ++j) tma_load_2d_2sm(sb + j * 8192, &tma_b, &full_bar[s], n0 + 64 * j, k0);
                    }
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== UMMA issuer (leader CTA only) =====================
        if (rank == 0) {
            // warp-uniform issue loop (descriptor arithmetic on the uniform datapath); one elected lane executes tcgen05.*
            const uint32_t leader = elect_one();
            constexpr uint32_t idesc = make_idesc_bf16(256, BN, A_MN, B_MN);
            int s = 0;
            uint32_t ph = 0;
            int as = 0;
            uint32_t aph = 0;
            WorkIter it(p, cluster_id, num_clusters, num_kb);
            WorkItem w;
            while (it.next(w)) {
                mbar_wait_cluster(&tmem_empty[as], aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = w.k0; kb < w.k1; ++kb) {
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
                    const uint32_t sb = sa + Cfg::A_BYTES;
                    const uint32_t acc0 = kb != w.k0;
                    if (leader) {
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            const uint64_t da = A_MN ? make_smem_desc_sw128(sa + k * (UMMA_K * 128), 8192, 1024)
                                                     : make_smem_desc_sw128(sa + k * (UMMA_K * 2), 16, 1024);
                            const uint64_t db = B_MN ? make_smem_desc_sw128(sb + k * (UMMA_K * 128), 8192, 1024)
                                                     : make_smem_desc_sw128(sb + k * (UMMA_K * 2), 16, 1024);
                            umma_ss_2sm(d_tmem, da, db, idesc, acc0 | (k != 0));
                        }
                        umma_commit_2sm_mc(&empty_bar[s], 0b11);   // frees the stage in BOTH CTAs
                    }
                    __syncwarp();
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                if (leader) umma_commit_2sm_mc(&tmem_full[as], 0b11);       // accumulator complete -> both epilogues
                __syncwarp();
                if (++as == 2) { as = 0; aph ^= 1; }
            }
        }
    } else {
        // ===================== epilogue warps (2..9), both CTAs =====================
        const int q = warp & 3;                                              // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;                                    // which half of the tile's columns it owns
        constexpr int EPI_THREADS = 32 * GEMM2_EPI_WARPS;
        constexpr int CH = BN / 64;                                          // 32-column chunks per warp
        const int et = threadIdx.x - 64;
        int as = 0;
        uint32_t aph = 0;
        EpiRegs e;
        e.N = p.N; e.ldc = p.ldc; e.ld_aux = p.ld_aux; e.C = reinterpret_cast<char*>(p.C); e.aux = reinterpret_cast<char*>(p.aux);
        e.bias = p.bias;
        const bool has_bias = (EPI != UD_EPI_F32 && EPI != UD_EPI_F32_ACC) && e.bias != nullptr;
        float alpha = 1.f;                     // UD_EPI_BF16_SCALED: C = bf16(bf16(acc) * alpha), alpha = *aux (1 / world size)
        if constexpr (EPI == UD_EPI_BF16_SCALED) alpha = *reinterpret_cast<const float*>(e.aux);
        e.alpha = alpha;
        float ssq = 0.f;                       // UD_EPI_F32 + aux: sum of squares of everything this thread stores (gradient norm)
        WorkIter it(p, cluster_id, num_clusters, num_kb);
        WorkItem w;
        while (it.next(w)) {
            int tm, tn;
            tile_coords(w.tile, p.num_m_tiles, p.num_n_tiles, tm, tn);
            const int m0 = tm * 256 + (int)rank * BM;
            const int n0 = tn * BN;
            const int c_lo = half * CH;
            const int row = m0 + q * 32 + lane;
            const bool row_ok = row < p.M;
            float* bias_s = bias_smem + as * BN;
            // ---- everything that does not need the accumulator happens BEFORE waiting for it (overlaps the main loop) ----
            // GELU-backward reads the saved pre-activations u[row, n0 + ...]: issued here, BEFORE waiting for the accumulator, as
            // coalesced loads (8 lanes x 16 B per row, 4 rows per instruction); transposed to row-per-thread through the
            // warp's shared-memory buffer at the point of use.
            // (one 64-column slab is in flight at a time: the next slab's loads are issued while the current one is processed —
            //  holding all of a 256-wide tile's pre-activations cost 64 registers and spilled, see -Xptxas -v)
            uint4 pre[8];
            bool aux_coal = false;
            const int rows_v = max(0, min(32, p.M - (m0 + q * 32)));
            auto prefetch_pre = [&](int sl) {
                const int uu = lane & 7, r0 = lane >> 3;
                const int col0 = n0 + (c_lo + 2 * sl) * 32;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rr = i * 4 + r0;
                    pre[i] = make_uint4(0u, 0u, 0u, 0u);
                    if (rr < rows_v && col0 + 64 <= e.N)
                        pre[i] = ldg_stream(reinterpret_cast<const uint4*>(
                            e.aux + ((long long)(m0 + q * 32 + rr) * e.ld_aux + col0) * 2 + uu * 16));
                }
            };
            if constexpr (EPI == UD_EPI_BF16_DGELU) {
                aux_coal = (e.ld_aux & 7) == 0;
                if (w.kind != 2 && aux_coal) prefetch_pre(0);
            }
            if (has_bias) {
                // this tile's bias, fp32, in shared memory (one barrier per work item: a buffer is rewritten two items later,
                // after everybody passed the barrier in between)
                if (w.kind != 2) {
#pragma unroll
                    for (int j = et; j < BN; j += EPI_THREADS) bias_s[j] = (n0 + j < e.N) ? __bfloat162float(e.bias[n0 + j]) : 0.f;
                }
                named_bar_sync(1, EPI_THREADS);
            }
            mbar_wait(&tmem_full[as], aph);
            tc_fence_after();
            const int trow = q * 32 + lane;                                  // row inside this CTA's 128-row half
            const uint32_t tacc = tmem_base + as * BN + ((uint32_t)(q * 32) << 16);
            if (w.kind == 2) {
                // ---- stream-K contributor: park the fp32 partial accumulator in this cluster's workspace slot ----
                float* ws = p.sk_ws + ((long long)cluster_id * 2 + rank) * (BN * 128);
#pragma unroll 1
                for (int c = c_lo; c < c_lo + CH; ++c) {
                    if (n0 + c * 32 >= e.N) break;
                    uint32_t r[32];
                    tmem_ld_32x32b_x32(tacc + c * 32, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) ws[(c * 32 + j) * 128 + trow] = __uint_as_float(r[j]);   // coalesced over lanes
                }
                tc_fence_before();
                __threadfence();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_cluster_relaxed(&tmem_empty[as], 0);
                    atomicAdd(p.sk_flags + w.sk_t, 1);
                }
            } else {
                int n_part = 0, first_c = 0;
                if (w.kind == 1) {
                    // ---- stream-K owner: wait for the other clusters' partials of this tile (they finished earlier or
                    //      at the same time: a contributor's segment is never preceded by a wait) ----
                    const long long it0 = (long long)w.sk_t * num_kb;
                    first_c = (int)(it0 / p.sk_w);
                    const int last_c = (int)((it0 + num_kb - 1) / p.sk_w);
                    n_part = last_c - first_c;
                    if (lane == 0) {
                        const int want = n_part * 2 * GEMM2_EPI_WARPS;
                        int seen;
                        do {
                            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(p.sk_flags + w.sk_t) : "memory");
                            if (seen < want) __nanosleep(64);
                        } while (seen < want);
                    }
                    __syncwarp();
                }
                auto add_partials = [&](uint32_t (&r)[32], int cc) {
                    for (int pc = 1; pc <= n_part; ++pc) {
                        const float* ws = p.sk_ws + ((long long)(first_c + pc) * 2 + rank) * (BN * 128) + (cc * 32) * 128 + trow;
#pragma unroll
                        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __ldcg(ws + j * 128));
                    }
                };
                uint8_t* sbuf = store_smem + (warp - 2) * 4096;
                const int row_base = m0 + q * 32;
                const int rows_ok = max(0, min(32, p.M - row_base));
                if constexpr (EPI == UD_EPI_F32 || EPI == UD_EPI_F32_ACC) {
                    // fp32 outputs: one 32-column chunk = 128 bytes per row = one slab
#pragma unroll
                    for (int c = 0; c < CH; ++c) {
                        const int cc = c_lo + c, col0 = n0 + cc * 32;
                        if (col0 < e.N) {
                            uint32_t r[32];
                            tmem_ld_32x32b_x32(tacc + cc * 32, r);
                            tmem_ld_wait();
                            add_partials(r, cc);
                            if constexpr (EPI == UD_EPI_F32) {
                                // rows >= M and columns >= N of the accumulator are exact zeros (TMA zero-fills out-of-bounds operands)
                                if (e.aux != nullptr) {
#pragma unroll
                                    for (int j = 0; j < 32; ++j) ssq = fmaf(__uint_as_float(r[j]), __uint_as_float(r[j]), ssq);
                                }
                            }
                            if (col0 + 32 <= e.N) {
                                uint4 u[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) u[j] = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
                                store_slab_coalesced<EPI == UD_EPI_F32_ACC>(sbuf, u, lane, e.C + ((long long)row_base * e.ldc + col0) * 4,
                                                                            e.ldc * 4, rows_ok);
                            } else {
                                epilogue_chunk2<EPI>(r, row, row_ok, col0, e, nullptr, nullptr);
                            }
                        }
                    }
                } else {
                    // bf16 outputs: two chunks = 64 columns = 128 bytes per row = one slab.  Only ONE 32-column chunk of the
                    // accumulator is live at a time (the fused GELU / GELU' epilogues need ~60 more registers than the plain one).
                    const bool coal = ((e.ldc & 7) == 0) && (EPI != UD_EPI_BF16_GELU || (e.ld_aux & 7) == 0);
#pragma unroll
                    for (int sl = 0; sl < CH / 2; ++sl) {
                        const int cc = c_lo + 2 * sl, col0 = n0 + cc * 32;
                        if (col0 >= e.N) break;
                        if (coal && col0 + 64 <= e.N) {
                            uint4 u[8];
                            if constexpr (EPI == UD_EPI_BF16_DGELU) {
                                if (aux_coal) {
                                    // transpose this slab's pre-activations to row-per-thread through the warp's smem buffer, then
                                    // put the NEXT slab's loads in flight
                                    const int uu = lane & 7, r0q = lane >> 3;
#pragma unroll
                                    for (int i = 0; i < 8; ++i) {
                                        const int rr = i * 4 + r0q;
                                        *reinterpret_cast<uint4*>(sbuf + rr * 128 + ((uu ^ (rr & 7)) << 4)) = pre[i];
                                    }
                                    __syncwarp();
                                    if (sl + 1 < CH / 2) prefetch_pre(sl + 1);
                                }
                            }
#pragma unroll
                            for (int hch = 0; hch < 2; ++hch) {
                                uint32_t r[32];
                                tmem_ld_32x32b_x32(tacc + (cc + hch) * 32, r);
                                uint4 xg[4];                                 // this row's 32 pre-activations of the chunk (GELU-backward)
                                if constexpr (EPI == UD_EPI_BF16_DGELU) {
                                    if (aux_coal) {
#pragma unroll
                                        for (int j = 0; j < 4; ++j)
                                            xg[j] = *reinterpret_cast<const uint4*>(sbuf + lane * 128 + (((4 * hch + j) ^ (lane & 7)) << 4));
                                    } else if (row_ok) {
                                        const uint4* xr = reinterpret_cast<const uint4*>(e.aux + ((long long)row * e.ld_aux + col0 + 32 * hch) * 2);
#pragma unroll
                                        for (int j = 0; j < 4; ++j) xg[j] = xr[j];
                                    }
                                }
                                tmem_ld_wait();
                                add_partials(r, cc + hch);
                                float v[32];
#pragma unroll
                                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                                if (has_bias) {
                                    const float* bs = bias_s + (cc + hch) * 32;
#pragma unroll
                                    for (int j = 0; j < 8; ++j) {
                                        const float4 b = *reinterpret_cast<const float4*>(bs + 4 * j);
                                        v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
                                    }
                                }
                                if constexpr (EPI == UD_EPI_BF16_SCALED) {
                                    // torch's bf16 compress hook: buffer.to(bf16).div_(world) — round, scale, round
#pragma unroll
                                    for (int j = 0; j < 32; ++j) v[j] = bf16_round(v[j]) * alpha;
                                }
                                if constexpr (EPI == UD_EPI_BF16_DGELU) {
#pragma unroll
                                    for (int j = 0; j < 4; ++j) {
                                        const uint4 x = xg[j];
                                        v[8 * j + 0] *= gelu_tanh_grad(bf16lo(x.x)); v[8 * j + 1] *= gelu_tanh_grad(bf16hi(x.x));
                                        v[8 * j + 2] *= gelu_tanh_grad(bf16lo(x.y)); v[8 * j + 3] *= gelu_tanh_grad(bf16hi(x.y));
                                        v[8 * j + 4] *= gelu_tanh_grad(bf16lo(x.z)); v[8 * j + 5] *= gelu_tanh_grad(bf16hi(x.z));
                                        v[8 * j + 6] *= gelu_tanh_grad(bf16lo(x.w)); v[8 * j + 7] *= gelu_tanh_grad(bf16hi(x.w));
                                    }
                                }
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    u[4 * hch + j] = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                                                pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
                            }
                            if constexpr (EPI == UD_EPI_BF16_DGELU) __syncwarp();      // every lane has read its pre-activation row
                            store_slab_coalesced<false>(sbuf, u, lane, e.C + ((long long)row_base * e.ldc + col0) * 2, e.ldc * 2, rows_ok);
                            if constexpr (EPI == UD_EPI_BF16_GELU) {
                                // g = gelu(u) in place (u has been staged to shared memory by the store above)
#pragma unroll
                                for (int j = 0; j < 8; ++j)
                                    u[j] = make_uint4(pack_bf16x2(gelu_tanh(bf16lo(u[j].x)), gelu_tanh(bf16hi(u[j].x))),
                                                      pack_bf16x2(gelu_tanh(bf16lo(u[j].y)), gelu_tanh(bf16hi(u[j].y))),
                                                      pack_bf16x2(gelu_tanh(bf16lo(u[j].z)), gelu_tanh(bf16hi(u[j].z))),
                                                      pack_bf16x2(gelu_tanh(bf16lo(u[j].w)), gelu_tanh(bf16hi(u[j].w))));
                                store_slab_coalesced<false>(sbuf, u, lane, e.aux + ((long long)row_base * e.ld_aux + col0) * 2, e.ld_aux * 2,
                                                            rows_ok);
                            }
                        } else {
                            if constexpr (EPI == UD_EPI_BF16_DGELU) {
                                if (aux_coal && sl + 1 < CH / 2) prefetch_pre(sl + 1);
                            }
                            uint32_t r[32];
                            tmem_ld_32x32b_x32(tacc + cc * 32, r);
                            tmem_ld_wait();
                            add_partials(r, cc);
                            epilogue_chunk2<EPI>(r, row, row_ok, col0, e, has_bias ? bias_s + cc * 32 : nullptr, nullptr);
                            if (col0 + 32 < e.N) {
                                tmem_ld_32x32b_x32(tacc + (cc + 1) * 32, r);
                                tmem_ld_wait();
                                add_partials(r, cc + 1);
                                epilogue_chunk2<EPI>(r, row, row_ok, col0 + 32, e, has_bias ? bias_s + (cc + 1) * 32 : nullptr, nullptr);
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_cluster_relaxed(&tmem_empty[as], 0);
                    if (w.kind == 1) {
                        // the last of the owner's epilogue warps re-arms the counters for the next launch
                        if (atomicAdd(p.sk_flags + p.sk_tiles + w.sk_t, 1) == 2 * GEMM2_EPI_WARPS - 1) {
                            p.sk_flags[w.sk_t] = 0;
                            p.sk_flags[p.sk_tiles + w.sk_t] = 0;
                        }
                    }
                }
            }
            if (++as == 2) { as = 0; aph ^= 1; }
        }
        if constexpr (EPI == UD_EPI_F32) {
            if (e.aux != nullptr) {                // one atomic per epilogue warp per launch
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) ssq += __shfl_xor_sync(0xffffffffu, ssq, o);
                if (lane == 0 && ssq != 0.f) atomicAdd(reinterpret_cast<float*>(e.aux), ssq);
            }
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm<Cfg::TMEM_COLS>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_encodeTiled>(p);
    });
    return fn;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                      uint32_t box_cols) {
    PFN_encodeTiled fn = get_encode_fn();
    if (!fn) { fprintf(stderr, "unidisc_b200: cuTensorMapEncodeTiled unavailable\n"); return -1; }
    if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld * 2) & 15)) {
        fprintf(stderr, "unidisc_b200: TMA needs 16-byte aligned base (%p) and row pitch (ld=%llu)\n", base, (unsigned long long)ld);
        return -2;
    }
    cuuint64_t gdim[2] = {cols, rows};
    cuuint64_t gstr[1] = {ld * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        fprintf(stderr, "unidisc_b200: cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu box=%ux%u\n", (int)r,
                (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows, box_cols);
        return -3;
    }
    return 0;
}

int make_tmap_3d_bf16(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1, uint64_t s2,
                      uint32_t b0, uint32_t b1, uint32_t b2) {
    PFN_encodeTiled fn = get_encode_fn();
    if (!fn) return -1;
    cuuint64_t gdim[3] = {d0, d1, d2};
    cuuint64_t gstr[2] = {s1 * 2, s2 * 2};
    cuuint32_t box[3] = {b0, b1, b2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        fprintf(stderr, "unidisc_b200: cuTensorMapEncodeTiled(3d) failed (%d)\n", (int)r);
        return -3;
    }
    return 0;
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    }
    return n;
}

// stream-K workspace: one slot per cluster + arrival counters, cached per stream (GEMMs on different streams may overlap)
struct SkWorkspace { float* ws = nullptr; int* flags = nullptr; };
static int get_sk_workspace(cudaStream_t stream, SkWorkspace& out) {
    static std::mutex mu;
    static std::unordered_map<cudaStream_t, SkWorkspace> pool;
    std::lock_guard<std::mutex> lock(mu);
    auto it = pool.find(stream);
    if (it == pool.end()) {
        SkWorkspace w;
        const int clusters = sm_count() / 2;
        UD_CUDA_CHECK(cudaMalloc(&w.ws, (size_t)clusters * 2 * 256 * 128 * sizeof(float)));
        UD_CUDA_CHECK(cudaMalloc(&w.flags, (size_t)2 * clusters * sizeof(int)));
        UD_CUDA_CHECK(cudaMemset(w.flags, 0, (size_t)2 * clusters * sizeof(int)));
        it = pool.emplace(stream, w).first;
    }
    out = it->second;
    return 0;
}

template <bool A_MN, bool B_MN, int BN, int EPI>
static int launch_gemm2(const CUtensorMap& ta, const CUtensorMap& tb, GemmParams p, cudaStream_t stream) {
    using Cfg = Gemm2Cfg<BN>;
    auto kern = gemm2_kernel<A_MN, B_MN, BN, EPI>;
    static bool attr_set = false;
    if (!attr_set) {
        UD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_set = true;
    }
    const int tiles = p.num_m_tiles * p.num_n_tiles;
    const int C = sm_count() / 2;
    const int num_kb = (p.K + BK - 1) / BK;
    const int rem = tiles % C;
    static const bool sk_off = getenv("UD_GEMM_NO_STREAMK") != nullptr;
    // stream-K the partial last wave when it would leave > 7 % of the clusters idle and there is enough K to share
    // (measured on B200, tools/kbench.py: K >= 6144 with >= 1 full wave gains 2-9 %; K = 2048 tails and the 64-tile wgrad
    //  lose 10-18 % to the fix-up latency and to the k-offsets of the clusters no longer sharing operand panels in L2)
    static const bool sk_all = getenv("UD_GEMM_STREAMK_ALL") != nullptr;
    const bool sk_shape = sk_all ? (num_kb >= 4 && (long long)rem * num_kb >= C) : (num_kb >= 64 && tiles >= C);
    const bool use_sk = !sk_off && rem > 0 && sk_shape && rem * 100 < C * 93;
    int clusters = tiles < C ? tiles : C;
    p.sk_first_tile = tiles; p.sk_tiles = 0; p.sk_w = 1; p.sk_ws = nullptr; p.sk_flags = nullptr;
    if (use_sk) {
        SkWorkspace w;
        if (int rc = get_sk_workspace(stream, w)) return rc;
        p.sk_first_tile = tiles - rem;
        p.sk_tiles = rem;
        p.sk_w = (int)(((long long)rem * num_kb + C - 1) / C);
        p.sk_ws = w.ws;
        p.sk_flags = w.flags;
        clusters = C;
    }
    kern<<<2 * clusters, GEMM2_THREADS, Cfg::SMEM_BYTES, stream>>>(ta, tb, p);
    UD_CUDA_CHECK(cudaGetLastError());
    return 0;
}

template <bool A_MN, bool B_MN, int BN>
static int dispatch_epi2(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t s) {
    switch (epi) {
        case UD_EPI_BF16: return launch_gemm2<A_MN, B_MN, BN, UD_EPI_BF16>(ta, tb, p, s);
        case UD_EPI_BF16_GELU: return launch_gemm2<A_MN, B_MN, BN, UD_EPI_BF16_GELU>(ta, tb, p, s);
        case UD_EPI_BF16_DGELU: return launch_gemm2<A_MN, B_MN, BN, UD_EPI_BF16_DGELU>(ta, tb, p, s);
        case UD_EPI_F32: return launch_gemm2<A_MN, B_MN, BN, UD_EPI_F32>(ta, tb, p, s);
        case UD_EPI_F32_ACC: return launch_gemm2<A_MN, B_MN, BN, UD_EPI_F32_ACC>(ta, tb, p, s);
        case UD_EPI_BF16_SCALED:
            if constexpr (A_MN && B_MN) return launch_gemm2<A_MN, B_MN, BN, UD_EPI_BF16_SCALED>(ta, tb, p, s);
            break;
    }
    return -4;
}

template <int BN>
static int dispatch_major2(int ta_, int tb_, int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                           cudaStream_t s) {
    if (!ta_ && !tb_) return dispatch_epi2<false, false, BN>(epi, ta, tb, p, s);
    if (!ta_ && tb_) return dispatch_epi2<false, true, BN>(epi, ta, tb, p, s);
    if (ta_ && tb_) return dispatch_epi2<true, true, BN>(epi, ta, tb, p, s);
    return -5;
}

}  // namespace ud

extern "C" int ud_gemm_bf16(int ta, int tb, int M, int N, int K, const void* A, long long lda, const void* B, long long ldb,
                            void* C, long long ldc, int epi, const void* bias, void* aux, long long ld_aux, int bn_hint,
                            void* stream) {
    using namespace ud;
    if (M <= 0 || N <= 0 || K <= 0) return 0;
    // bn_hint: 0 = auto, 128 / 256 = tile width (other bits are ignored: the single-CTA kernel family of round 1 is gone, the
    // CTA-pair kernel covers every M — rows beyond M are TMA zero-fill and never stored)
    int BN = bn_hint & 1023;
    constexpr int TM = 256;
    const int units = sm_count() / 2;
    if (BN != 128 && BN != 256) {
        // 256-wide tiles run ~1.6x faster per flop than 128-wide ones (operand traffic from L2 per flop, measured on B200:
        // ~1560 vs ~950 TFLOP/s), so only fall back to 128 when wave quantisation costs more than that.
        auto cost = [&](int bn) {
            long long t = (long long)((M + TM - 1) / TM) * ((N + bn - 1) / bn);
            long long w = (t + units - 1) / units;
            return (double)w * bn * (bn == 128 ? 1.6 : 1.0);
        };
        BN = (cost(256) <= cost(128)) ? 256 : 128;
    }
    CUtensorMap tmA, tmB;
    int rc;
    if (!ta) rc = make_tmap_2d_bf16(&tmA, A, M, K, lda, BM, 64);
    else rc = make_tmap_2d_bf16(&tmA, A, K, M, lda, 64, 64);
    if (rc) return rc;
    if (!tb) rc = make_tmap_2d_bf16(&tmB, B, N, K, ldb, BN / 2, 64);
    else rc = make_tmap_2d_bf16(&tmB, B, K, N, ldb, 64, 64);
    if (rc) return rc;
    if ((reinterpret_cast<uintptr_t>(C) & 15) || (ldc % 4) != 0) {
        fprintf(stderr, "unidisc_b200: GEMM output needs a 16-byte aligned base and ldc %% 4 == 0\n");
        return -6;
    }
    if (aux != nullptr && epi == UD_EPI_F32_ACC) {
        fprintf(stderr, "unidisc_b200: the fused sum of squares (aux with an fp32 epilogue) needs UD_EPI_F32\n");
        return -7;
    }
    if (epi == UD_EPI_BF16_SCALED && (aux == nullptr || !(ta && tb))) {
        fprintf(stderr, "unidisc_b200: UD_EPI_BF16_SCALED is the weight-gradient epilogue (ta = tb = 1) and needs aux = device fp32 scale\n");
        return -8;
    }
    GemmParams p;
    p.M = M; p.N = N; p.K = K;
    p.C = C; p.ldc = ldc;
    p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
    p.aux = aux; p.ld_aux = ld_aux;
    p.num_m_tiles = (M + TM - 1) / TM;
    p.num_n_tiles = (N + BN - 1) / BN;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (BN == 256) return dispatch_major2<256>(ta, tb, epi, tmA, tmB, p, s);
    return dispatch_major2<128>(ta, tb, epi, tmA, tmB, p, s);
}
