// Interleaved-batch preparation on the device (sm_100a) — SURVEY.md §8 a18 / (f)2.
//
// Replaces the reference's host-side Python loops over image / sample blocks (models/dit.py:122-191 add_img_data_to_blocks,
// add_txt_data_to_blocks, unidisc/utils/tensor_utils.py:4-44) with ONE kernel launch: from `modality` and `sample_ids`
// it derives, per token, which RoPE table row to use and the image ordinal that indexes `img_count_embedding`, and
// gathers the per-token cos/sin rows the attention pre-pass consumes.
//
//   image block  = maximal run of modality != 0 in a row (regardless of sample boundaries, like the reference);
//                  if its size is one of {256,1024,2304,4096} its tokens take row (pos - run start) of that size's 2-D
//                  table and ordinal = #earlier image blocks of the row whose first token has the same sample id;
//                  other sizes keep cos = sin = 0 and ordinal = -1 (no table, no count embedding).
//   text token   = modality == 0 inside a run of equal sample_id >= 0: row (pos - run start) of the 1-D table;
//                  pad runs (sample_id < 0) keep cos = sin = 0.
//
// One CTA per batch row; run starts are found with a block-wide max-scan, run sizes are published by each run's last
// token, ordinals by each run's first token.  Pure integer work + a coalesced row gather; HBM-trivial.
#include "common.cuh"
#include "unidisc_b200.h"

namespace ud {

// inclusive max-scan over the row: thread t owns elements [t*C, t*C+C).  v[] holds the per-element candidates (index or -1).
template <int MAXC>
UD_DEVINL void block_max_scan(int (&v)[MAXC], int C, int* warp_tot /*[32]*/) {
    int run = -1;
#pragma unroll
    for (int j = 0; j < MAXC; ++j) {
        if (j < C) { run = max(run, v[j]); v[j] = run; }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    int incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl = max(incl, n);
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < nwarps ? warp_tot[lane] : -1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w = max(w, n);
        }
        warp_tot[lane] = w;     // inclusive over warps
    }
    __syncthreads();
    int excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = -1;
    if (warp > 0) excl = max(excl, warp_tot[warp - 1]);
#pragma unroll
    for (int j = 0; j < MAXC; ++j)
        if (j < C) v[j] = max(v[j], excl);
    __syncthreads();
}

constexpr int kMaxC = 16;   // elements per thread: N <= 16 * 1024

__global__ void __launch_bounds__(1024)
interleaved_prep_kernel(const int64_t* __restrict__ modality, const int64_t* __restrict__ sample_ids, int N,
                        const float* __restrict__ cos_tab, const float* __restrict__ sin_tab, int hd2, int txt_off,
                        int off256, int off1024, int off2304, int off4096, float* __restrict__ cos_out,
                        float* __restrict__ sin_out, int* __restrict__ ordinal_out, int* __restrict__ scratch /*[B,4N]*/) {
    __shared__ int warp_tot[32];
    __shared__ int n_starts;
    const int b = blockIdx.x;
    const int64_t* mod = modality + (long long)b * N;
    const int64_t* sid = sample_ids + (long long)b * N;
    int* size_at = scratch + (long long)b * 4 * N;     // [N] run size, valid at image-run starts
    int* ord_at = size_at + N;                          // [N] ordinal, valid at image-run starts
    int* starts = ord_at + N;                           // [<=N] list of image-run starts (unordered)
    int* src_row = starts + N;                          // [N] table row to gather for the token, -1 = zeros
    const int C = (N + blockDim.x - 1) / blockDim.x;
    const int base = threadIdx.x * C;
    if (threadIdx.x == 0) n_starts = 0;
    __syncthreads();

    int img_start[kMaxC], sid_start[kMaxC];
#pragma unroll
    for (int j = 0; j < kMaxC; ++j) {
        img_start[j] = sid_start[j] = -1;
        const int i = base + j;
        if (j < C && i < N) {
            const bool m = mod[i] != 0;
            if (m && (i == 0 || mod[i - 1] == 0)) { img_start[j] = i; starts[atomicAdd(&n_starts, 1)] = i; }
            if (i == 0 || sid[i] != sid[i - 1]) sid_start[j] = i;
        }
    }
    block_max_scan<kMaxC>(img_start, C, warp_tot);
    block_max_scan<kMaxC>(sid_start, C, warp_tot);
    // last token of every image run publishes the run size at the run's start
#pragma unroll
    for (int j = 0; j < kMaxC; ++j) {
        const int i = base + j;
        if (j < C && i < N && mod[i] != 0 && (i == N - 1 || mod[i + 1] == 0)) size_at[img_start[j]] = i + 1 - img_start[j];
    }
    __syncthreads();
    // first token of every image run counts the earlier runs of the same packed sample
    const int ns = n_starts;
#pragma unroll
    for (int j = 0; j < kMaxC; ++j) {
        const int i = base + j;
        if (j < C && i < N && mod[i] != 0 && img_start[j] == i) {
            const int64_t my = sid[i];
            int cnt = 0;
            for (int k = 0; k < ns; ++k) {
                const int s = starts[k];
                cnt += (s < i && sid[s] == my) ? 1 : 0;
            }
            ord_at[i] = cnt;
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kMaxC; ++j) {
        const int i = base + j;
        if (j < C && i < N) {
            int row = -1, ord = -1;
            if (mod[i] != 0) {
                const int s = img_start[j], size = size_at[s];
                const int off = size == 256 ? off256 : size == 1024 ? off1024 : size == 2304 ? off2304 : size == 4096 ? off4096 : -1;
                if (off >= 0) { row = off + (i - s); ord = ord_at[s]; }
            } else {
                const int s = sid_start[j];
                if (sid[s] >= 0) row = txt_off + (i - s);
            }
            src_row[i] = row;
            ordinal_out[(long long)b * N + i] = ord;
        }
    }
    __syncthreads();
    // coalesced gather of the cos/sin rows (float4 per thread)
    const int v4 = hd2 / 4;
    for (int e = threadIdx.x; e < N * v4; e += blockDim.x) {
        const int i = e / v4, c = (e % v4) * 4;
        const int row = src_row[i];
        float4 cv = make_float4(0.f, 0.f, 0.f, 0.f), sv = cv;
        if (row >= 0) {
            cv = *reinterpret_cast<const float4*>(cos_tab + (long long)row * hd2 + c);
            sv = *reinterpret_cast<const float4*>(sin_tab + (long long)row * hd2 + c);
        }
        *reinterpret_cast<float4*>(cos_out + ((long long)b * N + i) * hd2 + c) = cv;
        *reinterpret_cast<float4*>(sin_out + ((long long)b * N + i) * hd2 + c) = sv;
    }
}

}  // namespace ud

using namespace ud;

extern "C" int ud_interleaved_prep(const int64_t* modality, const int64_t* sample_ids, int B, int N, const float* cos_tab,
                                   const float* sin_tab, int hd2, int txt_off, int off256, int off1024, int off2304, int off4096,
                                   float* cos_out, float* sin_out, int* ordinal_out, int* scratch, void* stream) {
    if (B <= 0 || N <= 0) return 0;
    if (N > kMaxC * 1024) { fprintf(stderr, "unidisc_b200: interleaved_prep supports N <= %d (got %d)\n", kMaxC * 1024, N); return -1; }
    if (hd2 % 4 != 0) { fprintf(stderr, "unidisc_b200: interleaved_prep needs head_dim %% 8 == 0\n"); return -1; }
    int threads = 1024;
    while (threads > 32 && threads / 2 >= N) threads /= 2;
    interleaved_prep_kernel<<<B, threads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        modality, sample_ids, N, cos_tab, sin_tab, hd2, txt_off, off256, off1024, off2304, off4096, cos_out, sin_out, ordinal_out, scratch);
    UD_CUDA_CHECK(cudaGetLastError());
    return 0;
}
