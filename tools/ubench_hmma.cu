// Issue rate of the warp-level (legacy) tensor-core path on sm_100a: mma.sync m16n8k8 tf32 and m16n8k16 bf16, by the number of
// independent accumulator chains per warp and warps per SM sub-partition.  Built here (nvcc -arch=sm_100a), run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_hmma tools/ubench_hmma.cu && tools/ubench_hmma
// Answers what bounds k_prequant_prep (vq_prequant.cu): 3 x 2*T*C*D tf32 flop through this path.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int kChains, bool kBf16>
__global__ void k_mma(float* out, int iters, long long* cycles) {
    float d[kChains][4];
#pragma unroll
    for (int c = 0; c < kChains; ++c)
#pragma unroll
        for (int q = 0; q < 4; ++q) d[c][q] = 0.f;
    uint32_t a0 = threadIdx.x, a1 = threadIdx.x * 3, a2 = threadIdx.x * 5, a3 = threadIdx.x * 7, b0 = 11, b1 = 13;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < kChains; ++c) {
            if (kBf16)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < kChains; ++c) s += d[c][0] + d[c][1] + d[c][2] + d[c][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int kChains, bool kBf16>
void run(int warps_per_block, int sms) {
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * 1024 * 1024); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    k_mma<kChains, kBf16><<<sms, warps_per_block * 32>>>(out, 16, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_mma<kChains, kBf16><<<sms, warps_per_block * 32>>>(out, iters, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double mmas_per_smsp = (double)iters * kChains * warps_per_block / 4.0;
    const double flop = 2.0 * 16 * 8 * (kBf16 ? 16 : 8) * (double)iters * kChains * warps_per_block * sms;
    printf("%s chains/warp %d warps/SM %2d: %6.2f cycles per MMA and sub-partition, %7.1f TFLOP/s over %d SMs\n",
           kBf16 ? "bf16 m16n8k16" : "tf32 m16n8k8 ", kChains, warps_per_block, (double)c / mmas_per_smsp, flop / (ms * 1e-3) / 1e12, sms);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    run<1, false>(4, sms); run<4, false>(4, sms); run<8, false>(4, sms); run<8, false>(8, sms); run<8, false>(16, sms);
    run<1, true>(4, sms); run<8, true>(4, sms); run<8, true>(8, sms); run<8, true>(16, sms);
    return 0;
}
