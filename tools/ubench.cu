// Micro-benchmarks of the sm_100a primitives the attention kernels depend on (dev tool, run on the GPU box):
//   A  tcgen05.ld bandwidth per SM (4 / 8 warps)         B  MUFU.EX2 throughput per SM
//   C  tcgen05.mma issue -> commit -> mbarrier wake-up latency (1 vs 8 MMAs)      D  mbarrier arrive -> waiter wake-up
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I unidisc_b200/csrc -I include -o tools/ubench tools/ubench.cu
#include <cstdio>
#include <vector>
#include "common.cuh"
using namespace ud;

__global__ void k_tmem_ld(long long* out, int iters) {
    __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc<512>(&tptr);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t t = tptr + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        uint32_t a[32], b[32], c[32], d[32];
        const uint32_t col = (uint32_t)((it * 128 + (warp >> 2) * 64) & 255);
        tmem_ld_32x32b_x32(t + col, a);
        tmem_ld_32x32b_x32(t + col + 32, b);
        tmem_ld_32x32b_x32(t + col + 64, c);
        tmem_ld_32x32b_x32(t + col + 96, d);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) acc ^= a[i] ^ b[i] ^ c[i] ^ d[i];
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) out[blockIdx.x] = 0;
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tptr); }
}

__global__ void k_mufu(long long* out, float* sink, int iters) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 0.1f, x2 = x0 + 0.2f, x3 = x0 + 0.3f;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x0));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x1));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x2));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x3));
        x0 -= 1.0f; x1 -= 1.0f; x2 -= 1.0f; x3 -= 1.0f;
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
}

// one thread issues `nmma` MMAs (128 x N x 16, operands = zeroed smem) then commits and spins on the barrier
template <int N>
__global__ void k_mma_latency(long long* out, int nmma, int reps) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint32_t tptr;
    __shared__ uint64_t bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<512>(&tptr);
    fence_proxy_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (warp == 1 && lane == 0) {
        constexpr uint32_t idesc = make_idesc_bf16(128, N, false, false);
        const uint32_t sa = smem_u32(smem), sb = sa + 16384;
        long long tot = 0;
        for (int r = 0; r < reps; ++r) {
            const long long t0 = clock64();
            for (int k = 0; k < nmma; ++k)
                umma_ss(tptr, make_smem_desc_sw128(sa + (k & 3) * 32, 16, 1024), make_smem_desc_sw128(sb + (k & 3) * 32, 16, 1024), idesc, k != 0);
            umma_commit(&bar);
            while (!mbar_try_wait(&bar, r & 1)) {}
            tot += clock64() - t0;
        }
        out[blockIdx.x] = tot / reps;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tptr); }
}


// MMA-side cost of one attention-backward sub-tile: the tensor pipe's instruction mix issued back to back by one thread,
// no softmax in between.  VAR 0: current dK/dV kernel (64-query sub-tile): 2 x [8 SS 128x64x16] + 2 x [4 TS 128x128x16, B MN-major]
// VAR 1: 128-query sub-tile: 2 x [8 SS 128x128x16] + 2 x [8 TS 128x128x16]   VAR 2: SS part of VAR 0   VAR 3: TS part of VAR 0
// VAR 4: SS part of VAR 1   VAR 5: TS part of VAR 1
template <int VAR>
__global__ void k_mma_seq(long long* out, int reps) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint32_t tptr;
    __shared__ uint64_t bar;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<512>(&tptr);
    fence_proxy_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (warp == 1 && lane == 0) {
        constexpr uint32_t id64 = make_idesc_bf16(128, 64, false, false), id128 = make_idesc_bf16(128, 128, false, false);
        constexpr uint32_t idacc = make_idesc_bf16(128, 128, false, true);
        const uint32_t fa = smem_u32(smem), fb = fa + 32768, sa = fa + 65536, sb = sa + 32768;
        auto kmaj = [](uint32_t base, int ks, int rows) { return make_smem_desc_sw128(base + (ks >> 2) * (rows * 128) + (ks & 3) * 32, 16, 1024); };
        auto mnmaj = [](uint32_t base, int ks, int rows) { return make_smem_desc_sw128(base + ks * 2048, rows * 128, 1024); };
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const uint32_t tS = tptr + (r & 1) * 128, tD = tS + 64;
            if (VAR == 0 || VAR == 2) {
                for (int ks = 0; ks < 8; ++ks) umma_ss(tS, kmaj(fa, ks, 128), kmaj(sa, ks, 64), id64, ks != 0);
                for (int ks = 0; ks < 8; ++ks) umma_ss(tD, kmaj(fb, ks, 128), kmaj(sb, ks, 64), id64, ks != 0);
            }
            if (VAR == 0 || VAR == 3) {
                for (int ks = 0; ks < 4; ++ks) umma_ts(tptr + 256, tS + ks * 8, mnmaj(sb, ks, 64), idacc, 1);
                for (int ks = 0; ks < 4; ++ks) umma_ts(tptr + 384, tD + ks * 8, mnmaj(sa, ks, 64), idacc, 1);
            }
            if (VAR == 1 || VAR == 4) {
                for (int ks = 0; ks < 8; ++ks) umma_ss(tptr, kmaj(fa, ks, 128), kmaj(sa, ks, 128), id128, ks != 0);
                for (int ks = 0; ks < 8; ++ks) umma_ss(tptr + 128, kmaj(fb, ks, 128), kmaj(sb, ks, 128), id128, ks != 0);
            }
            if (VAR == 6 || VAR == 7) {   // scores with the fixed operand (A) in TMEM: only the 64-row B tile is read from smem
                for (int ks = 0; ks < 8; ++ks) umma_ts(tS, tptr + 384 + ks * 8, kmaj(sa, ks, 64), id64, ks != 0);
                for (int ks = 0; ks < 8; ++ks) umma_ts(tD, tptr + 448 + ks * 8, kmaj(sb, ks, 64), id64, ks != 0);
            }
            if (VAR == 7) {
                for (int ks = 0; ks < 4; ++ks) umma_ts(tptr + 256, tD + ks * 8, mnmaj(sa, ks, 64), idacc, 1);
            }
            if (VAR == 1 || VAR == 5) {
                for (int ks = 0; ks < 8; ++ks) umma_ts(tptr + 256, tptr + ks * 8, mnmaj(sb, ks, 128), idacc, 1);
                for (int ks = 0; ks < 8; ++ks) umma_ts(tptr + 384, tptr + 128 + ks * 8, mnmaj(sa, ks, 128), idacc, 1);
            }
        }
        umma_commit(&bar);
        while (!mbar_try_wait(&bar, 0)) {}
        out[blockIdx.x] = (clock64() - t0) / reps;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tptr); }
}

// warp 0 lane 0 arrives, warp 1 lane 0 waits: round-trip ping-pong latency
__global__ void k_mbar_pingpong(long long* out, int reps) {
    __shared__ uint64_t b0, b1;
    if (threadIdx.x == 0) { mbar_init(&b0, 1); mbar_init(&b1, 1); fence_barrier_init(); }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            mbar_arrive(&b0);
            while (!mbar_try_wait(&b1, r & 1)) {}
        }
        out[blockIdx.x] = (clock64() - t0) / reps;
    } else if (warp == 1 && lane == 0) {
        for (int r = 0; r < reps; ++r) {
            while (!mbar_try_wait(&b0, r & 1)) {}
            mbar_arrive(&b1);
        }
    }
}

static double avg(const std::vector<long long>& v) { double s = 0; for (auto x : v) s += x; return s / v.size(); }

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long* d; cudaMalloc(&d, sizeof(long long) * sms);
    float* sink; cudaMalloc(&sink, sizeof(float) * sms * 1024);
    std::vector<long long> h(sms);
    for (int threads : {128, 256}) {
        const int iters = 2000;
        k_tmem_ld<<<sms, threads>>>(d, iters); cudaDeviceSynchronize();
        k_tmem_ld<<<sms, threads>>>(d, iters); cudaDeviceSynchronize();
        cudaMemcpy(h.data(), d, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
        const double bytes = (double)iters * (threads / 32) * 4 * 32 * 32 * 4;
        printf("tmem_ld  %d warps/SM: %.1f cycles, %.1f B/clk/SM (err=%s)\n", threads / 32, avg(h), bytes / avg(h), cudaGetErrorString(cudaGetLastError()));
    }
    for (int threads : {128, 256, 512}) {
        const int iters = 4000;
        k_mufu<<<sms, threads>>>(d, sink, iters); cudaDeviceSynchronize();
        cudaMemcpy(h.data(), d, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
        printf("mufu.ex2 %d warps/SM: %.2f ex2/clk/SM\n", threads / 32, (double)iters * 4 * threads / avg(h));
    }
    cudaFuncSetAttribute(k_mma_latency<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(k_mma_latency<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(k_mma_latency<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int nm : {1, 4, 8, 16, 32}) {
        k_mma_latency<64><<<sms, 64, 64 * 1024>>>(d, nm, 200); cudaDeviceSynchronize();
        cudaMemcpy(h.data(), d, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
        const double a64 = avg(h);
        k_mma_latency<128><<<sms, 64, 64 * 1024>>>(d, nm, 200); cudaDeviceSynchronize();
        cudaMemcpy(h.data(), d, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
        const double a128 = avg(h);
        k_mma_latency<256><<<sms, 64, 64 * 1024>>>(d, nm, 200); cudaDeviceSynchronize();
        cudaMemcpy(h.data(), d, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
        printf("mma issue->commit->wake, %2d MMAs (128xNx16): N=64 %.0f  N=128 %.0f  N=256 %.0f cycles (err=%s)\n", nm, a64, a128, avg(h), cudaGetErrorString(cudaGetLastError()));
    }
    {
        auto run = [&](auto kern, const char* name) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 164 * 1024);
            kern<<<sms, 64, 164 * 1024>>>(d, 40); cudaDeviceSynchronize();
            kern<<<sms, 64, 164 * 1024>>>(d, 40); cudaDeviceSynchronize();
            cudaMemcpy(h.data(), d, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
            printf("mma seq %-56s %.0f cycles/rep (err=%s)\n", name, avg(h), cudaGetErrorString(cudaGetLastError()));
        };
        run(k_mma_seq<0>, "v2 sub-tile (64q): 16 SS N=64 + 8 TS N=128");
        run(k_mma_seq<2>, "  SS part: 16 x 128x64x16");
        run(k_mma_seq<3>, "  TS part: 8 x 128x128x16 (B MN-major)");
        run(k_mma_seq<1>, "128q sub-tile: 16 SS N=128 + 16 TS N=128");
        run(k_mma_seq<4>, "  SS part: 16 x 128x128x16");
        run(k_mma_seq<5>, "  TS part: 16 x 128x128x16 (B MN-major)");
        run(k_mma_seq<6>, "dQ scores, A in TMEM: 16 TS 128x64x16 (B K-major)");
        run(k_mma_seq<7>, "dQ sub-tile, A in TMEM: 16 TS N=64 + 4 TS N=128");
    }
    k_mbar_pingpong<<<sms, 64>>>(d, 1000); cudaDeviceSynchronize();
    cudaMemcpy(h.data(), d, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    printf("mbarrier ping-pong round trip: %.0f cycles (err=%s)\n", avg(h), cudaGetErrorString(cudaGetLastError()));
    return 0;
}
