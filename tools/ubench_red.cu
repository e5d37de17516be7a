// Micro-benchmark: throughput of 64-bit integer reductions (RED.ADD.64) into a K x D table, the access pattern of
// the codebook-gradient segment sums when they are accumulated straight from the token pass (no bucketing).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_red tools/ubench_red.cu && ./ubench_red
// Patterns (8 lanes per token row of D = 32 unless noted):
//   0  lane l owns elements 4l .. 4l+3        (4 REDs, each touching 8 sectors per row)
//   1  lane l owns elements l, l+8, l+16, l+24 (4 REDs, each touching 2 sectors per row)
//   2  32 lanes per row, one element each      (1 RED, 8 sectors, whole 256-byte row)
//   3  pattern 0 with the stream of row reads only (no RED): the floor
// Code distributions: uniform over K, 128 hot codes, 1 hot code.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

template <int PAT>
__global__ void __launch_bounds__(256) k_red(const int* __restrict__ idx, const float4* __restrict__ zn, int T,
                                             unsigned long long* __restrict__ table, float* sink) {
    const int lane = threadIdx.x & 31;
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int stride = gridDim.x * blockDim.x;
    float acc = 0.f;
    if (PAT == 2) {
        for (int row = gtid >> 5; row < T; row += stride >> 5) {
            const int k = __ldg(idx + row);
            const float v = __ldg(reinterpret_cast<const float*>(zn) + (int64_t)row * 32 + lane);
            atomicAdd(table + (int64_t)k * 32 + lane, (unsigned long long)__float2ll_rn(v * 1073741824.f));
        }
    } else {
        const int m = lane & 7;
        for (int row = gtid >> 3; row < T; row += stride >> 3) {
            const int k = __ldg(idx + row);
            const float4 v = __ldg(zn + (int64_t)row * 8 + m);
            if (PAT == 3) { acc += v.x + v.y + v.z + v.w + k; continue; }
            float e[4] = {v.x, v.y, v.z, v.w};
            if (PAT == 1) {
                // 4x4 transpose inside each group of 4 lanes: afterwards lane l (l&3 = j) holds element 4*l' + j ... the
                // benchmark only needs the address pattern, so the values are not moved
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    atomicAdd(table + (int64_t)k * 32 + m + 8 * i, (unsigned long long)__float2ll_rn(e[i] * 1073741824.f));
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    atomicAdd(table + (int64_t)k * 32 + 4 * m + i, (unsigned long long)__float2ll_rn(e[i] * 1073741824.f));
            }
        }
    }
    if (acc == 12345.678f) *sink = acc;
}

int main() {
    const int T = 262144, K = 8192, D = 32;
    int* idx; float4* zn; unsigned long long* table; float* sink;
    cudaMalloc(&idx, sizeof(int) * T);
    cudaMalloc(&zn, sizeof(float) * T * D);
    cudaMalloc(&table, sizeof(unsigned long long) * K * D);
    cudaMalloc(&sink, 4);
    cudaMemset(zn, 0x3c, sizeof(float) * T * D);
    char* flush; cudaMalloc(&flush, 256u << 20);
    std::vector<int> h(T);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const char* dist_name[3] = {"uniform", "128 hot codes", "1 hot code"};
    for (int dist = 0; dist < 3; ++dist) {
        srand(1);
        for (int t = 0; t < T; ++t) h[t] = dist == 0 ? rand() % K : (dist == 1 ? (rand() % 128) * 64 : 77);
        cudaMemcpy(idx, h.data(), sizeof(int) * T, cudaMemcpyHostToDevice);
        for (int pat = 0; pat < 4; ++pat) {
            for (int blocks : {148 * 4, 148 * 8, 148 * 16}) {
                float best = 1e9f;
                for (int rep = 0; rep < 5; ++rep) {
                    cudaMemsetAsync(flush, rep, 256u << 20);
                    cudaMemsetAsync(table, 0, sizeof(unsigned long long) * K * D);
                    cudaEventRecord(a);
                    if (pat == 0) k_red<0><<<blocks, 256>>>(idx, zn, T, table, sink);
                    if (pat == 1) k_red<1><<<blocks, 256>>>(idx, zn, T, table, sink);
                    if (pat == 2) k_red<2><<<blocks, 256>>>(idx, zn, T, table, sink);
                    if (pat == 3) k_red<3><<<blocks, 256>>>(idx, zn, T, table, sink);
                    cudaEventRecord(b);
                    cudaEventSynchronize(b);
                    float ms; cudaEventElapsedTime(&ms, a, b);
                    if (ms < best) best = ms;
                }
                printf("%-14s pattern %d blocks %5d : %8.2f us\n", dist_name[dist], pat, blocks, best * 1e3f);
            }
        }
    }
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
