// Shared device-side helpers for the sm_100a kernels: mbarrier, TMA, tcgen05 (UMMA/TMEM) PTX wrappers.
// Written against the PTX ISA as shipped with CUDA 12.9; compile with -gencode arch=compute_100a,code=sm_100a.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define UD_DEVINL __device__ __forceinline__

namespace ud {

// ------------------------------------------------------------------------------------------------
// small utilities
// ------------------------------------------------------------------------------------------------
UD_DEVINL uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

UD_DEVINL float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

UD_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
UD_DEVINL float bf16lo(uint32_t v) { return __uint_as_float(v << 16); }
UD_DEVINL float bf16hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

UD_DEVINL float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
UD_DEVINL float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// fast tanh: 1 - 2/(exp(2x)+1); abs error ~1e-6, far below the bf16 rounding that follows every use.
UD_DEVINL float fast_tanh(float x) {
    float e = __expf(2.0f * x);
    return 1.0f - __fdividef(2.0f, e + 1.0f);
}
// single-MUFU tanh (max rel. error 2^-11, well below the bf16 rounding that follows every use in the GEMM epilogues)
UD_DEVINL float tanh_mufu(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// GELU(tanh) and derivative — nn.GELU(approximate="tanh") (reference models/dit.py:918)
UD_DEVINL float gelu_tanh(float x) {
    const float k = 0.7978845608028654f, kc = 0.7978845608028654f * 0.044715f;
    const float x2 = x * x;
    const float t = tanh_mufu(x * fmaf(kc, x2, k));
    const float hx = 0.5f * x;
    return fmaf(hx, t, hx);
}
UD_DEVINL float gelu_tanh_grad(float x) {
    const float k = 0.7978845608028654f, kc = 0.7978845608028654f * 0.044715f;
    const float x2 = x * x;
    const float t = tanh_mufu(x * fmaf(kc, x2, k));
    const float dt = fmaf(-t, t, 1.0f) * fmaf(3.0f * kc, x2, k);     // (1 - t^2) * k (1 + 3c x^2)
    return fmaf(0.5f * x, dt, fmaf(0.5f, t, 0.5f));
}

// ------------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------------
UD_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
UD_DEVINL void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
UD_DEVINL void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

UD_DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
UD_DEVINL void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
UD_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (kernel aborts with an error) instead of hanging the GPU.
UD_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {  // ~2 s at 1.9 GHz
            printf("ud: mbarrier wait timeout block=(%d,%d,%d) thread=%d bar=%p parity=%u\n", blockIdx.x, blockIdx.y,
                   blockIdx.z, threadIdx.x, (void*)bar, parity);
            __trap();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) — 2D / 3D tiled loads into shared memory, completion on an mbarrier
// ------------------------------------------------------------------------------------------------
UD_DEVINL void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
UD_DEVINL void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
UD_DEVINL void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
UD_DEVINL void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
UD_DEVINL void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
UD_DEVINL void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
UD_DEVINL void tma_store_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// ------------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, UMMA issue/commit, TMEM <-> register moves, fences
// ------------------------------------------------------------------------------------------------
template <uint32_t NCOLS>
UD_DEVINL void tmem_alloc(uint32_t* smem_dst) {  // whole warp, .sync.aligned
    static_assert(NCOLS == 32 || NCOLS == 64 || NCOLS == 128 || NCOLS == 256 || NCOLS == 512, "power of 2 >= 32");
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(NCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t NCOLS>
UD_DEVINL void tmem_dealloc(uint32_t taddr) {  // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(NCOLS) : "memory");
}
UD_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
UD_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// One elected lane of a fully converged warp.  The MMA-issuing warp runs its loop warp-uniformly and only the tcgen05
// instructions are guarded by this predicate, so descriptor / address arithmetic stays on the uniform datapath (inside an
// `if (lane == 0)` region every operand is a per-thread register and each tcgen05.mma costs ~15 instructions + R2UR moves).
UD_DEVINL uint32_t elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred;
}

// D[tmem] (+)= A[smem desc] * B[smem desc]   (kind::f16: bf16/fp16 inputs, fp32 accumulate)
UD_DEVINL void umma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]
UD_DEVINL void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on an mbarrier when all tcgen05 ops previously issued by THIS thread have completed.
UD_DEVINL void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// ---- cta_group::2 (CTA pair on one TPC): cluster helpers, 2-SM alloc / MMA / commit / TMA ----
UD_DEVINL uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
UD_DEVINL void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <uint32_t NCOLS>
UD_DEVINL void tmem_alloc_2sm(uint32_t* smem_dst) {  // one warp in EACH CTA of the pair, same smem offset
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(NCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t NCOLS>
UD_DEVINL void tmem_dealloc_2sm(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(NCOLS) : "memory");
}
// D[tmem of both CTAs] (+)= A * B with M = 256 split across the pair; issued by the leader CTA only
UD_DEVINL void umma_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on the mbarrier at the same smem offset in every CTA of `cta_mask` once prior MMAs of this thread retire
UD_DEVINL void umma_commit_2sm_mc(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(cta_mask)
                 : "memory");
}
// TMA load into THIS CTA's smem whose completion bytes are credited to the LEADER CTA's mbarrier (peer bit cleared)
UD_DEVINL void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}
// arrive on the mbarrier at the same smem offset in CTA `cta` of the cluster
UD_DEVINL void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
// same, relaxed: no memory ordering of the caller's prior global stores is needed (and none is paid for: the .release form
// compiles to MEMBAR.ALL + ERRBAR, which waits for every outstanding epilogue store).  Used for TMEM hand-offs, where the
// tcgen05.wait::ld + tcgen05.fence::before_thread_sync pair already orders the TMEM reads.
UD_DEVINL void mbar_arrive_cluster_relaxed(uint64_t* bar, uint32_t cta) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
UD_DEVINL bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
UD_DEVINL void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {
            printf("ud: cluster mbarrier wait timeout block=%d thread=%d parity=%u\n", blockIdx.x, threadIdx.x, parity);
            __trap();
        }
    }
}

// UMMA instruction descriptor (kind::f16), bf16 x bf16 -> fp32. Bit layout: cute/arch/mma_sm100_desc.hpp InstrDescriptor.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
    return (1u << 4)                        // c_format = F32
           | (1u << 7)                      // a_format = BF16
           | (1u << 10)                     // b_format = BF16
           | ((a_mn_major ? 1u : 0u) << 15) // a_major
           | ((b_mn_major ? 1u : 0u) << 16) // b_major
           | ((uint32_t)(N >> 3) << 17)     // n_dim
           | ((uint32_t)(M >> 4) << 24);    // m_dim
}

// Shared-memory matrix descriptor, SWIZZLE_128B, sm_100 format (version=1). Addresses/offsets in 16-byte units.
//  K-major operand tile  [rows][64 bf16]  (128-byte rows, 8-row swizzle atoms 1024 B apart): LBO unused, SBO = 1024 B.
//  MN-major operand tile [k rows][64 bf16 along MN] stacked per 64-wide MN chunk:              LBO = chunk stride, SBO = 1024 B.
UD_DEVINL uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
    d |= (uint64_t)2 << 61;  // layout_type = SWIZZLE_128B
    return d;
}

// TMEM -> registers: 32 lanes x 32-bit, N consecutive columns; thread i of the warp gets lane (base_lane + i).
UD_DEVINL void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
UD_DEVINL void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
UD_DEVINL void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
UD_DEVINL void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
UD_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
UD_DEVINL void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// named barrier over a subset of the CTA's threads
UD_DEVINL void named_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// streaming (no L1 allocate) 16-byte global accesses for read-once / write-once data
UD_DEVINL uint4 ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (counter-based RNG for the non-parity "fast" sampling mode)
// ------------------------------------------------------------------------------------------------
UD_DEVINL uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0; key.y += W1;
    }
    return ctr;
}
// uniform in [0,1) with 24 random bits for element index `idx`
UD_DEVINL float philox_uniform(uint64_t seed, uint64_t offset, uint64_t idx) {
    const uint64_t blk = idx >> 2;
    uint4 r = philox4x32_10(make_uint4((uint32_t)blk, (uint32_t)(blk >> 32), (uint32_t)offset, (uint32_t)(offset >> 32)),
                            make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const uint32_t w = (idx & 3) == 0 ? r.x : (idx & 3) == 1 ? r.y : (idx & 3) == 2 ? r.z : r.w;
    return (float)(w >> 8) * (1.0f / 16777216.0f);
}

// dropout keep-scales for the 4 consecutive hidden columns [4*cg, 4*cg+4) of token row `row`.  One Philox call yields 4 x 32
// random bits = the 16-bit draws of a column group for TWO rows (2r, 2r+1): the row kernels that process row pairs call
// dropout_bits once per pair (the 10 Philox rounds were ~10 % of their instructions).  keep with probability 1-p
// (u16 >= p * 65536), kept values are scaled by 1/(1-p) like F.dropout.
UD_DEVINL uint4 dropout_bits(uint64_t seed, uint64_t offset, uint32_t row_pair, uint32_t cg) {
    return philox4x32_10(make_uint4(row_pair, cg, (uint32_t)offset, (uint32_t)(offset >> 32)),
                         make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
}
UD_DEVINL void dropout_pick(const uint4& r, uint32_t odd_row, uint32_t thresh, float inv_keep, float (&ks)[4]) {
    const uint32_t sh = odd_row * 16u, t16 = thresh >> 16;
    ks[0] = ((r.x >> sh) & 0xffffu) >= t16 ? inv_keep : 0.f;
    ks[1] = ((r.y >> sh) & 0xffffu) >= t16 ? inv_keep : 0.f;
    ks[2] = ((r.z >> sh) & 0xffffu) >= t16 ? inv_keep : 0.f;
    ks[3] = ((r.w >> sh) & 0xffffu) >= t16 ? inv_keep : 0.f;
}
UD_DEVINL void dropout_scales4(uint64_t seed, uint64_t offset, uint32_t row, uint32_t cg, uint32_t thresh, float inv_keep,
                               float (&ks)[4]) {
    dropout_pick(dropout_bits(seed, offset, row >> 1, cg), row & 1u, thresh, inv_keep, ks);
}
inline uint32_t dropout_thresh(float p) {
    double t = (double)p * 4294967296.0;
    return t >= 4294967295.0 ? 0xffffffffu : (uint32_t)t;
}

}  // namespace ud

// ------------------------------------------------------------------------------------------------
// host-side helpers
// ------------------------------------------------------------------------------------------------
#define UD_CUDA_CHECK(expr)                                                                              \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            fprintf(stderr, "unidisc_b200: CUDA error %s at %s:%d: %s\n", cudaGetErrorName(_e), __FILE__, \
                    __LINE__, cudaGetErrorString(_e));                                                   \
            return (int)_e;                                                                              \
        }                                                                                                \
    } while (0)

namespace ud {
// Build a 2-D bf16 tiled tensor map (row-major [rows, cols], leading dimension ld elements), SWIZZLE_128B.
// box = {box_cols (<=64 for 128B swizzle), box_rows}. Returns 0 on success.
int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                      uint32_t box_cols);
// 3-D variant: [d2, d1, d0] with strides (in elements) s2, s1, 1.
int make_tmap_3d_bf16(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1, uint64_t s2,
                      uint32_t b0, uint32_t b1, uint32_t b2);
int sm_count();
}  // namespace ud
