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
};

template <int BN>
struct GemmCfg {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (BN == 256) ? 4 : 6;
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

// Epilogue for one 32-column chunk of one accumulator row held in registers (shared by the 1-CTA and 2-CTA kernels).
template <int EPI>
UD_DEVINL void epilogue_chunk(const uint32_t (&r)[32], int row, bool row_ok, int col0, const GemmParams& p) {
    const bool full = (col0 + 32 <= p.N);
    if constexpr (EPI == UD_EPI_F32 || EPI == UD_EPI_F32_ACC) {
        float* cp = reinterpret_cast<float*>(p.C) + (long long)row * p.ldc + col0;
        if (row_ok) {
            if (full) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 v = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                           __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
                    if constexpr (EPI == UD_EPI_F32_ACC) {
                        float4 o = *reinterpret_cast<float4*>(cp + 4 * j);
                        v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
                    }
                    *reinterpret_cast<float4*>(cp + 4 * j) = v;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (col0 + j < p.N) {
                        float v = __uint_as_float(r[j]);
                        if constexpr (EPI == UD_EPI_F32_ACC) v += cp[j];
                        cp[j] = v;
                    }
                }
            }
        }
    } else {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (p.bias != nullptr) {
            if (full) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint4 b = __ldg(reinterpret_cast<const uint4*>(p.bias + col0) + j);
                    v[8 * j + 0] += bf16lo(b.x); v[8 * j + 1] += bf16hi(b.x);
                    v[8 * j + 2] += bf16lo(b.y); v[8 * j + 3] += bf16hi(b.y);
                    v[8 * j + 4] += bf16lo(b.z); v[8 * j + 5] += bf16hi(b.z);
                    v[8 * j + 6] += bf16lo(b.w); v[8 * j + 7] += bf16hi(b.w);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (col0 + j < p.N) v[j] += __bfloat162float(p.bias[col0 + j]);
            }
        }
        __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(p.C) + (long long)row * p.ldc + col0;
        __nv_bfloat16* xp = reinterpret_cast<__nv_bfloat16*>(p.aux) + (long long)row * p.ld_aux + col0;
        if (row_ok) {
            if constexpr (EPI == UD_EPI_BF16_DGELU) {
                // C = acc * gelu'(u), u = aux (bf16 pre-activation saved by the forward)
                if (full) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint4 u = ldg_stream(reinterpret_cast<const uint4*>(xp) + j);
                        v[8 * j + 0] *= gelu_tanh_grad(bf16lo(u.x)); v[8 * j + 1] *= gelu_tanh_grad(bf16hi(u.x));
                        v[8 * j + 2] *= gelu_tanh_grad(bf16lo(u.y)); v[8 * j + 3] *= gelu_tanh_grad(bf16hi(u.y));
                        v[8 * j + 4] *= gelu_tanh_grad(bf16lo(u.z)); v[8 * j + 5] *= gelu_tanh_grad(bf16hi(u.z));
                        v[8 * j + 6] *= gelu_tanh_grad(bf16lo(u.w)); v[8 * j + 7] *= gelu_tanh_grad(bf16hi(u.w));
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (col0 + j < p.N) v[j] *= gelu_tanh_grad(__bfloat162float(xp[j]));
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
                    for (int j = 0; j < 16; ++j)
                        g[j] = pack_bf16x2(gelu_tanh(bf16lo(o[j])), gelu_tanh(bf16hi(o[j])));
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *(reinterpret_cast<uint4*>(xp) + j) = make_uint4(g[4 * j], g[4 * j + 1], g[4 * j + 2], g[4 * j + 3]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (col0 + j < p.N) {
                        __nv_bfloat16 ub = __float2bfloat16_rn(v[j]);
                        cp[j] = ub;
                        if constexpr (EPI == UD_EPI_BF16_GELU) xp[j] = __float2bfloat16_rn(gelu_tanh(__bfloat162float(ub)));
                    }
                }
            }
        }
    }
}

template <bool A_MN, bool B_MN, int BN, int EPI>
__global__ void __launch_bounds__(192, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const GemmParams p) {
    using Cfg = GemmCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full = empty_bar + STAGES;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = p.num_m_tiles * p.num_n_tiles;
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
            mbar_init(&tmem_empty[s], 4);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_ptr_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int tm, tn;
                tile_coords(tile, p.num_m_tiles, p.num_n_tiles, tm, tn);
                const int m0 = tm * BM;
                const int n0 = tn * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
                    uint8_t* sb = sa + Cfg::A_BYTES;
                    mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
                    const int k0 = kb * BK;
                    if constexpr (!A_MN) {
                        tma_load_2d(sa, &tma_a, &full_bar[s], k0, m0);
                    } else {
#pragma unroll
                        for (int j = 0; j < BM / 64; ++j) tma_load_2d(sa + j * 8192, &tma_a, &full_bar[s], m0 + 64 * j, k0);
                    }
                    if constexpr (!B_MN) {
                        tma_load_2d(sb, &tma_b, &full_bar[s], k0, n0);
                    } else {
#pragma unroll
                        for (int j = 0; j < BN / 64; ++j) tma_load_2d(sb + j * 8192, &tma_b, &full_bar[s], n0 + 64 * j, k0);
                    }
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== UMMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN, B_MN);
            int s = 0;
            uint32_t ph = 0;
            int as = 0;
            uint32_t aph = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tmem_empty[as], aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
                    const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t da = A_MN ? make_smem_desc_sw128(sa + k * (UMMA_K * 128), 8192, 1024)
                                                 : make_smem_desc_sw128(sa + k * (UMMA_K * 2), 16, 1024);
                        const uint64_t db = B_MN ? make_smem_desc_sw128(sb + k * (UMMA_K * 128), 8192, 1024)
                                                 : make_smem_desc_sw128(sb + k * (UMMA_K * 2), 16, 1024);
                        umma_ss(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[s]);  // frees this smem stage when the MMAs above retire
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                umma_commit(&tmem_full[as]);     // accumulator complete -> epilogue
                if (++as == 2) { as = 0; aph ^= 1; }
            }
        }
    } else {
        // ===================== epilogue warps (2..5) =====================
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        int as = 0;
        uint32_t aph = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            int tm, tn;
            tile_coords(tile, p.num_m_tiles, p.num_n_tiles, tm, tn);
            const int m0 = tm * BM;
            const int n0 = tn * BN;
            mbar_wait(&tmem_full[as], aph);
            tc_fence_after();
            const int row = m0 + q * 32 + lane;
            const bool row_ok = row < p.M;
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                const int col0 = n0 + c * 32;
                if (col0 >= p.N) break;  // warp-uniform
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_base + as * BN + c * 32 + ((uint32_t)(q * 32) << 16), r);
                tmem_ld_wait();
                epilogue_chunk<EPI>(r, row, row_ok, col0, p);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[as]);
            if (++as == 2) { as = 0; aph ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
}


// ------------------------------------------------------------------------------------------------
// 2-CTA variant (cta_group::2): a CTA pair on one TPC computes a 256 x BN tile.  Each CTA stages its own 128 rows of A
// and HALF of the B tile (BN/2 rows), the leader's single thread issues M=256 UMMAs that read both CTAs' shared memory,
// and each CTA's TMEM receives its 128 accumulator rows.  Per flop this halves the B-operand traffic from L2, which is
// what bounds the 1-CTA kernel (128x256: 0.0117 B/flop ~ 20 TB/s at peak vs ~12 TB/s of L2).
// ------------------------------------------------------------------------------------------------
template <int BN>
struct Gemm2Cfg {
    static constexpr int A_BYTES = BM * BK * 2;            // 128 rows of A per CTA
    static constexpr int B_BYTES = (BN / 2) * BK * 2;      // half of the B tile per CTA
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = (BN == 256) ? 6 : 8;
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

template <bool A_MN, bool B_MN, int BN, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1)
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

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;
    const int num_tiles = p.num_m_tiles * p.num_n_tiles;   // tiles of 256 x BN
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
            mbar_init(&tmem_empty[s], 8);     // 4 epilogue warps in each of the two CTAs (used in the leader only)
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
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                int tm, tn;
                tile_coords(tile, p.num_m_tiles, p.num_n_tiles, tm, tn);
                const int m0 = tm * 256 + (int)rank * BM;
                const int n0 = tn * BN + (int)rank * (BN / 2);
                for (int kb = 0; kb < num_kb; ++kb) {
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
        if (rank == 0 && lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(256, BN, A_MN, B_MN);
            int s = 0;
            uint32_t ph = 0;
            int as = 0;
            uint32_t aph = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                mbar_wait_cluster(&tmem_empty[as], aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + s * Cfg::STAGE_BYTES);
                    const uint32_t sb = sa + Cfg::A_BYTES;
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t da = A_MN ? make_smem_desc_sw128(sa + k * (UMMA_K * 128), 8192, 1024)
                                                 : make_smem_desc_sw128(sa + k * (UMMA_K * 2), 16, 1024);
                        const uint64_t db = B_MN ? make_smem_desc_sw128(sb + k * (UMMA_K * 128), 8192, 1024)
                                                 : make_smem_desc_sw128(sb + k * (UMMA_K * 2), 16, 1024);
                        umma_ss_2sm(d_tmem, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit_2sm_mc(&empty_bar[s], 0b11);   // frees the stage in BOTH CTAs
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                umma_commit_2sm_mc(&tmem_full[as], 0b11);       // accumulator complete -> both epilogues
                if (++as == 2) { as = 0; aph ^= 1; }
            }
        }
    } else {
        // ===================== epilogue warps (2..5), both CTAs =====================
        const int q = warp & 3;
        int as = 0;
        uint32_t aph = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
            int tm, tn;
            tile_coords(tile, p.num_m_tiles, p.num_n_tiles, tm, tn);
            const int m0 = tm * 256 + (int)rank * BM;
            const int n0 = tn * BN;
            mbar_wait(&tmem_full[as], aph);
            tc_fence_after();
            const int row = m0 + q * 32 + lane;
            const bool row_ok = row < p.M;
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                const int col0 = n0 + c * 32;
                if (col0 >= p.N) break;
                uint32_t r[32];
                tmem_ld_32x32b_x32(tmem_base + as * BN + c * 32 + ((uint32_t)(q * 32) << 16), r);
                tmem_ld_wait();
                epilogue_chunk<EPI>(r, row, row_ok, col0, p);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&tmem_empty[as], 0);
            if (++as == 2) { as = 0; aph ^= 1; }
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

template <bool A_MN, bool B_MN, int BN, int EPI>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream) {
    using Cfg = GemmCfg<BN>;
    auto kern = gemm_kernel<A_MN, B_MN, BN, EPI>;
    static bool attr_set = false;
    if (!attr_set) {
        UD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_set = true;
    }
    int tiles = p.num_m_tiles * p.num_n_tiles;
    int grid = tiles < sm_count() ? tiles : sm_count();
    kern<<<grid, 192, Cfg::SMEM_BYTES, stream>>>(ta, tb, p);
    UD_CUDA_CHECK(cudaGetLastError());
    return 0;
}

template <bool A_MN, bool B_MN, int BN>
static int dispatch_epi(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t s) {
    switch (epi) {
        case UD_EPI_BF16: return launch_gemm<A_MN, B_MN, BN, UD_EPI_BF16>(ta, tb, p, s);
        case UD_EPI_BF16_GELU: return launch_gemm<A_MN, B_MN, BN, UD_EPI_BF16_GELU>(ta, tb, p, s);
        case UD_EPI_BF16_DGELU: return launch_gemm<A_MN, B_MN, BN, UD_EPI_BF16_DGELU>(ta, tb, p, s);
        case UD_EPI_F32: return launch_gemm<A_MN, B_MN, BN, UD_EPI_F32>(ta, tb, p, s);
        case UD_EPI_F32_ACC: return launch_gemm<A_MN, B_MN, BN, UD_EPI_F32_ACC>(ta, tb, p, s);
    }
    fprintf(stderr, "unidisc_b200: unknown GEMM epilogue %d\n", epi);
    return -4;
}

template <int BN>
static int dispatch_major(int ta_, int tb_, int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                          cudaStream_t s) {
    if (!ta_ && !tb_) return dispatch_epi<false, false, BN>(epi, ta, tb, p, s);
    if (!ta_ && tb_) return dispatch_epi<false, true, BN>(epi, ta, tb, p, s);
    if (ta_ && tb_) return dispatch_epi<true, true, BN>(epi, ta, tb, p, s);
    fprintf(stderr, "unidisc_b200: GEMM layout (ta=1,tb=0) is not used by the DiT path and not built\n");
    return -5;
}


template <bool A_MN, bool B_MN, int BN, int EPI>
static int launch_gemm2(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream) {
    using Cfg = Gemm2Cfg<BN>;
    auto kern = gemm2_kernel<A_MN, B_MN, BN, EPI>;
    static bool attr_set = false;
    if (!attr_set) {
        UD_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_set = true;
    }
    int tiles = p.num_m_tiles * p.num_n_tiles;
    int clusters = sm_count() / 2;
    if (tiles < clusters) clusters = tiles;
    kern<<<2 * clusters, 192, Cfg::SMEM_BYTES, stream>>>(ta, tb, p);
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
    // bn_hint: 0 = auto, 128 / 256 = tile width; +1024 forces the single-CTA kernel (A/B testing, tiny problems)
    const bool force_1cta = (bn_hint & 1024) != 0;
    int BN = bn_hint & 1023;
    const bool two_cta = !force_1cta && M > 128;
    const int TM = two_cta ? 256 : BM;
    const int units = two_cta ? sm_count() / 2 : sm_count();
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
    const int b_rows = two_cta ? BN / 2 : BN;
    if (!tb) rc = make_tmap_2d_bf16(&tmB, B, N, K, ldb, b_rows, 64);
    else rc = make_tmap_2d_bf16(&tmB, B, K, N, ldb, 64, 64);
    if (rc) return rc;
    if ((reinterpret_cast<uintptr_t>(C) & 15) || (ldc % 4) != 0) {
        fprintf(stderr, "unidisc_b200: GEMM output needs a 16-byte aligned base and ldc %% 4 == 0\n");
        return -6;
    }
    GemmParams p;
    p.M = M; p.N = N; p.K = K;
    p.C = C; p.ldc = ldc;
    p.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
    p.aux = aux; p.ld_aux = ld_aux;
    p.num_m_tiles = (M + TM - 1) / TM;
    p.num_n_tiles = (N + BN - 1) / BN;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (two_cta) {
        if (BN == 256) return dispatch_major2<256>(ta, tb, epi, tmA, tmB, p, s);
        return dispatch_major2<128>(ta, tb, epi, tmA, tmB, p, s);
    }
    if (BN == 256) return dispatch_major<256>(ta, tb, epi, tmA, tmB, p, s);
    return dispatch_major<128>(ta, tb, epi, tmA, tmB, p, s);
}
