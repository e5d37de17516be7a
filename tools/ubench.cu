// Micro-benchmarks that size the tensor-core epilogue: issue rate of FMNMX / FMNMX3 / FSEL / LOP3 and
// TMEM -> register load throughput (tcgen05.ld 32x32b.x32), per SM, for 1..4 warps per scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench tools/ubench.cu && ./ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float max3(float a, float b, float c) {
    float r; asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;
}
template <int OP>
__global__ void k_alu(float* out, long long* cyc, int iters, float seed) {
    float r[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = seed + j + threadIdx.x;
    float a = seed * 3.f, b = seed * 5.f;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            if (OP == 0) r[j] = fmaxf(r[j], a + j);                 // FMNMX (+FADD feeding it)
            if (OP == 1) r[j] = max3(r[j], a, b);                    // FMNMX3
            if (OP == 2) { asm volatile("max.f32 %0, %0, %1;" : "+f"(r[j]) : "f"(a)); }   // FMNMX only
            if (OP == 3) { asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(*(unsigned*)&r[j]) : "r"(0xFFFFFFE0u), "r"(j)); }
            if (OP == 4) { asm volatile("{.reg .pred p; setp.ne.b32 p, %2, 0; selp.f32 %0, %1, %0, p;}" : "+f"(r[j]) : "f"(a), "r"(i & 1)); }
            if (OP == 5) r[j] = fmaf(r[j], a, b);                    // FFMA reference
            if (OP == 6) { unsigned& u = *(unsigned*)&r[j]; asm volatile("{.reg .b32 t; max.s16x2 t, %0, %1; max.s16x2 %0, t, %2;}" : "+r"(u) : "r"(__float_as_uint(a)), "r"(__float_as_uint(b))); }   // VIMNMX3.S16x2
            if (OP == 7) { unsigned& u = *(unsigned*)&r[j]; asm volatile("max.s16x2 %0, %0, %1;" : "+r"(u) : "r"(__float_as_uint(a))); }   // VIMNMX.S16x2
            if (OP == 8) { unsigned& u = *(unsigned*)&r[j]; asm volatile("max.f16x2 %0, %0, %1;" : "+r"(u) : "r"(__float_as_uint(a))); }   // HMNMX2
            if (OP == 9) { unsigned& u = *(unsigned*)&r[j]; asm volatile("fma.rn.relu.f16x2 %0, %0, %1, %2;" : "+r"(u) : "r"(__float_as_uint(a)), "r"(__float_as_uint(b))); }   // HFMA2.RELU
            if (OP == 10) { unsigned& u = *(unsigned*)&r[j]; asm volatile("add.rn.f16x2 %0, %0, %1;" : "+r"(u) : "r"(__float_as_uint(a))); }   // HADD2
            if (OP == 11) { unsigned& u = *(unsigned*)&r[j]; asm volatile("{.reg .b32 t; max.s32 t, %0, %1; max.s32 %0, t, %2;}" : "+r"(u) : "r"(__float_as_uint(a)), "r"(__float_as_uint(b))); }   // VIMNMX3 (s32)
        }
        a += 1.f; b -= 1.f;
    }
    long long t1 = clock64();
    float s = 0; 
#pragma unroll
    for (int j = 0; j < 32; ++j) s += r[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void k_ldtm(float* out, long long* cyc, int iters) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
    float acc = 0.f;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint32_t u[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
                  "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
                  "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
                  "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
                : "r"(base + (uint32_t)(((i & 3) * 4 + c) * 32)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += __uint_as_float(u[0] ^ u[31]);
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

template <int OP> void run_alu(const char* name, float* out, long long* cyc) {
    for (int warps : {4, 8, 16}) {
        const int iters = 2000;
        k_alu<OP><<<148, warps * 32>>>(out, cyc, iters, 1.5f);
        cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        const double instr = (double)iters * 32 * warps;              // warp-instructions per SM
        printf("%-14s warps/SM=%2d  cycles=%lld  warp-instr/clk/SM=%.3f (lanes/clk/SM=%.1f)\n", name, warps, c,
               instr / c, instr * 32 / c);
    }
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    run_alu<2>("FMNMX", out, cyc);
    run_alu<1>("FMNMX3", out, cyc);
    run_alu<3>("LOP3", out, cyc);
    run_alu<4>("SELP", out, cyc);
    run_alu<5>("FFMA", out, cyc);
    run_alu<6>("VIMNMX3.S16x2", out, cyc);
    run_alu<7>("VIMNMX.S16x2", out, cyc);
    run_alu<8>("HMNMX2", out, cyc);
    run_alu<9>("HFMA2.RELU", out, cyc);
    run_alu<10>("HADD2", out, cyc);
    run_alu<11>("VIMNMX3.S32", out, cyc);
    for (int warps : {4, 8, 16}) {
        const int iters = 2000;
        k_ldtm<<<148, warps * 32>>>(out, cyc, iters);
        cudaError_t e = cudaDeviceSynchronize();
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        const double bytes = (double)iters * 4 * warps * 32 * 32 * 4;
        printf("LDTM.x32 warps/SM=%2d  cycles=%lld  bytes/clk/SM=%.1f  (%s)\n", warps, c, bytes / c, cudaGetErrorString(e));
    }
    return 0;
}
