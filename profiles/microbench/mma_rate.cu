// Issue-rate microbenchmark of tcgen05.mma kind::f16 (M128 x N x K16) on sm_100a for the operand forms the recurrences use:
//   0  A from tensor memory (TS), B shared memory, no-swizzle K-major core matrices   (k_gru_tc / k_gru_fx: W_hh, W_ih hi)
//   1  A shared memory no-swizzle core matrices (SS), B as in 0                       (W_lo of the H = 512 recurrence)
//   2  A shared memory 128-byte swizzle (SS), B as in 0
//   3  A and B shared memory 128-byte swizzle                                         (k_dwpw_bx / k_gemm form)
//   4  TS with B in 128-byte swizzle
// build + run on the GPU box:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I deepfilternet_b200/csrc \
//                              profiles/microbench/mma_rate.cu -o /tmp/mma_rate && /tmp/mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "dfb_ptx.cuh"
using namespace dfb;

template <int FL, int N>
__global__ void __launch_bounds__(128, 1) k_rate(int reps, long long *out) {
    extern __shared__ __align__(1024) unsigned char raw[];
    unsigned char *base = reinterpret_cast<unsigned char *>(((uintptr_t)raw + 1023) & ~uintptr_t(1023));
    unsigned char *sa = base, *sb = base + 65536;   // A: 128 rows x 256 K bf16 = 64 KB, B: N rows x 256 K x (hi|lo) <= 64 KB
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 131072 / 16; i += 128) reinterpret_cast<uint4 *>(base)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(&tmem_base, 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;
    if (warp == 1) {
        constexpr uint32_t idesc = umma_idesc_bf16(128, N);
        constexpr uint32_t kLbo = (N / 8) * 256, kSbo = 256;
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
        const uint32_t d = tmem_u + 384;
        const uint64_t b_il = umma_desc_interleave(smem_u32(sb), kLbo, kSbo);
        const uint64_t b_sw = umma_desc_sw128(smem_u32(sb));
        const uint64_t a_il = umma_desc_interleave(smem_u32(sa), 16 * 128, 128);
        long long t0 = clock64();
        for (int r = 0; r < reps; r++) {
#pragma unroll
            for (int ks = 0; ks < 16; ks++) {
                const uint64_t bi = b_il + (uint64_t)((ks * 2 * kLbo) >> 4);
                const uint64_t bs = umma_desc_sw128(smem_u32(sb) + (ks / 4) * (N * 128)) + 2 * (ks % 4);
                const uint64_t as = umma_desc_sw128(smem_u32(sa) + (ks / 4) * 16384) + 2 * (ks % 4);
                if (FL == 0) umma_bf16_ts_elect(d, tmem_u + ks * 8, bi, idesc, 1u);
                if (FL == 1) umma_bf16_ss_elect(d, a_il + (uint64_t)((ks * 2 * 16 * 128) >> 4), bi, idesc, 1u);
                if (FL == 2) umma_bf16_ss_elect(d, as, bi, idesc, 1u);
                if (FL == 3) umma_bf16_ss_elect(d, as, bs, idesc, 1u);
                if (FL == 4) umma_bf16_ts_elect(d, tmem_u + ks * 8, bs, idesc, 1u);
            }
        }
        long long t1 = clock64();
        umma_commit_elect(&bar);
        mbar_wait(&bar, 0);
        long long t2 = clock64();
        if (lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
        (void)b_sw;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

template <int FL, int N>
void run(long long *d_out) {
    const int reps = 64, smem = 131072 + 2048;
    cudaFuncSetAttribute(k_rate<FL, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long h[2];
    for (int it = 0; it < 2; it++) {
        k_rate<FL, N><<<1, 128, smem>>>(reps, d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("form %d N %d: %s\n", FL, N, cudaGetErrorString(e)); return; }
    }
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("form %d  N %2d : issue %7.1f cycles / MMA, complete %7.1f cycles / MMA  (%d MMAs)\n", FL, N, (double)h[0] / (reps * 16),
           (double)h[1] / (reps * 16), reps * 16);
}

int main() {
    long long *d_out;
    cudaMalloc(&d_out, 16);
    run<0, 16>(d_out); run<0, 32>(d_out); run<0, 48>(d_out); run<0, 64>(d_out);
    run<1, 16>(d_out); run<1, 32>(d_out); run<1, 64>(d_out);
    run<2, 16>(d_out); run<2, 32>(d_out); run<2, 64>(d_out);
    run<3, 16>(d_out); run<3, 32>(d_out); run<3, 64>(d_out);
    run<4, 16>(d_out); run<4, 32>(d_out); run<4, 64>(d_out);
    return 0;
}
