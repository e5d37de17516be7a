// Micro-benchmark of the tcgen05.mma issue path at codebook_dim = 32 shapes: cycles per M = 128, K = 16 MMA for
// N = 64 / 128 / 256 with both operands in shared memory (SS) or A in tensor memory (TS), fp16 and fp32
// accumulators, and the latencies that bound the accumulator ring of the nearest-code filter:
// MMA issue -> commit -> mbarrier visible, and tcgen05.ld (packed 16-bit) issue -> data in registers.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_mma tools/ubench_mma.cu && tools/ubench_mma
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {      // K-major, SWIZZLE_64B, 64-byte rows
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;
    return d;
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}

// mode 0: SS, 1: TS.  n: MMA N.  c_f32: accumulator format.  Result: cycles for `iters` MMAs issued back to back by one
// thread (accumulators rotate over the TMEM columns, B tiles over 4 shared-memory tiles, A over its two K = 16 halves).
__global__ void __launch_bounds__(128, 1) k_mma(long long* out, int mode, int n, int c_f32, int iters, int concurrent_lds) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar;
    const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
    uint8_t* smem = smem_raw + pad;
    for (int i = threadIdx.x; i < (8192 + 4 * 16384) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;  // 1.0h
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const uint32_t a_smem = smem_u32(smem), b_smem = a_smem + 8192;
    const uint32_t idesc = ((uint32_t)(c_f32 ? 1 : 0) << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const int n_acc = mode == 1 ? (448 / n) : (512 / n);           // TS: the last 64 columns hold A
    long long t0 = 0, t1 = 0;
    // the issue path as in the filter kernel: the whole warp loops converged, one elected lane issues, the warp index is
    // provably uniform (shuffled) so that descriptors live in the uniform datapath
    const int uwarp = __shfl_sync(0xffffffffu, warp, 0);
    if (uwarp == 0) {
        t0 = clock64();
        for (int i = 0; i < iters; i += 8) {
            uint32_t pred = 0;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
            if (pred) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t d = tmem + (uint32_t)(((i >> 1) + (u >> 1)) % n_acc) * n;
                    const uint64_t b = umma_desc(b_smem + ((u >> 1) & 3) * 16384 + (u & 1) * 32);
                    if (mode == 0) mma_ss(d, umma_desc(a_smem + (u & 1) * 32), b, idesc, u & 1);
                    else mma_ts(d, tmem + 448 + (u & 1) * 8, b, idesc, u & 1);
                }
            }
            __syncwarp();
        }
        if (threadIdx.x == 0) {
            commit(smem_u32(&bar));
            mbar_wait(smem_u32(&bar), 0);
            t1 = clock64();
            out[0] = t1 - t0;
        }
    } else if (concurrent_lds && warp >= 1) {
        // shared-memory read traffic next to the MMAs (does the operand fetch share the LDS crossbar?)
        uint32_t acc = 0;
        const uint32_t base = a_smem + (threadIdx.x & 31) * 16;
        for (int i = 0; i < iters * 8; ++i) {
            uint32_t x, y, z, w;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(base + ((i & 63) << 9)));
            acc += x ^ y ^ z ^ w;
        }
        if (acc == 0x12345) out[1] = acc;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// `issuers` warps issue `iters` MMAs each (own accumulator columns, shared operands); block time / total MMAs
__global__ void __launch_bounds__(256, 1) k_mma_multi(long long* out, int n, int iters, int issuers, int mode, int distinct, int ldtm) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar[4];
    __shared__ long long t_end[4];
    const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
    uint8_t* smem = smem_raw + pad;
    for (int i = threadIdx.x; i < (4 * 8192 + 16 * 8192) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const int uwarp = __shfl_sync(0xffffffffu, warp, 0);
    // distinct: every issuer has its own A tile (8 KB) and four own B tiles (8 KB each), as in the filter kernel where the
    // operands of consecutive MMAs never coincide; else all issuers read the same bytes
    const uint32_t a_smem = smem_u32(smem) + (distinct ? uwarp : 0) * 8192, b_smem = smem_u32(smem) + 4 * 8192 + (distinct ? uwarp : 0) * 4 * 8192;
    const uint32_t idesc = ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const long long t0 = clock64();
    if (uwarp < issuers) {
        const uint32_t d = tmem + (uint32_t)uwarp * (mode ? 96 : 128);      // TS: columns 448.. hold A
        for (int i = 0; i < iters; i += 8) {
            uint32_t pred = 0;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xFFFFFFFF;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
            if (pred) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint64_t b = umma_desc(b_smem + ((u >> 1) & 3) * 8192 + (u & 1) * 32);
                    if (mode == 0) mma_ss(d, umma_desc(a_smem + (u & 1) * 32), b, idesc, u & 1);
                    else mma_ts(d, tmem + 448 + (uint32_t)uwarp * 16 + (u & 1) * 8, b, idesc, u & 1);
                }
            }
            __syncwarp();
        }
        if ((threadIdx.x & 31) == 0) {
            commit(smem_u32(&bar[uwarp]));
            mbar_wait(smem_u32(&bar[uwarp]), 0);
            t_end[uwarp] = clock64();
        }
    } else if (ldtm && uwarp >= 4) {
        // the epilogue's traffic next to the MMAs: warps 4-7 keep reading accumulator tiles out of tensor memory
        // (x64.pack::16b, the filter kernel's load), about one 128-column tile per MMA pair
        const uint32_t taddr = tmem + ((uint32_t)((uwarp & 3) * 32) << 16);
        uint32_t acc = 0;
        for (int i = 0; i < iters / 2 * issuers; ++i) {
            uint32_t v[64];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x64.pack::16b.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
                "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
                "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                  "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                  "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                  "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]),
                  "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]),
                  "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]),
                  "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]),
                  "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
                : "r"(taddr + (uint32_t)((i & 3) * 128)));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            acc += v[0] ^ v[63];
        }
        if (acc == 0x12345) out[2] = acc;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        long long m = 0;
        for (int i = 0; i < issuers; ++i) m = t_end[i] > m ? t_end[i] : m;
        out[0] = m - t0;
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// latencies: (a) one N-wide stage (2 MMAs, K = 32) issue -> commit -> mbarrier visible; (b) tcgen05.ld of that stage
__global__ void __launch_bounds__(128, 1) k_lat(long long* out, int n, int reps) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar;
    const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
    uint8_t* smem = smem_raw + pad;
    for (int i = threadIdx.x; i < (8192 + 16384) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    const uint32_t a_smem = smem_u32(smem), b_smem = a_smem + 8192;
    const uint32_t idesc = ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    long long mma_lat = 0, ld_lat = 0;
    for (int r = 0; r < reps; ++r) {
        if (threadIdx.x == 0) {
            const long long t0 = clock64();
            mma_ss(tmem, umma_desc(a_smem), umma_desc(b_smem), idesc, 0);
            mma_ss(tmem, umma_desc(a_smem + 32), umma_desc(b_smem + 32), idesc, 1);
            commit(smem_u32(&bar));
            mbar_wait(smem_u32(&bar), r & 1);
            mma_lat += clock64() - t0;
        }
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        {
            uint32_t v[32];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
            const long long t0 = clock64();
            for (int c = 0; c < n / 64; ++c) {
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                      "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                      "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                      "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr + (uint32_t)(c * 64)));
            }
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const long long t1 = clock64();
            if (threadIdx.x == 0) { ld_lat += t1 - t0; if (v[0] == 0xdeadbeef) out[3] = v[5]; }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = mma_lat / reps; out[1] = ld_lat / reps; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
    long long* out;
    cudaMalloc(&out, 64);
    const int smem = 8192 + 4 * 16384 + 1024;
    cudaFuncSetAttribute(k_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k_lat, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 4096;
    for (int lds = 0; lds < 2; ++lds)
        for (int mode = 0; mode < 2; ++mode)
            for (int c_f32 = 0; c_f32 < 2; ++c_f32)
                for (int n : {64, 128, 256}) {
                    k_mma<<<148, 128, smem>>>(out, mode, n, c_f32, iters, lds);
                    cudaError_t e = cudaDeviceSynchronize();
                    long long c = 0;
                    cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
                    printf("%s acc=%s N=%3d %s: %7.1f cycles / MMA (M=128, K=16)   nominal %d   (%s)\n", mode ? "TS" : "SS",
                           c_f32 ? "f32" : "f16", n, lds ? "+LDS traffic" : "            ", (double)c / iters, n / 2, cudaGetErrorString(e));
                }
    // several issuer warps, each on its own accumulator columns: does the fixed per-MMA cost overlap between threads?
    cudaFuncSetAttribute(k_mma_multi, cudaFuncAttributeMaxDynamicSharedMemorySize, 20 * 8192 + 1024);
    for (int mode = 0; mode < 2; ++mode)
        for (int distinct = 0; distinct < 2; ++distinct)
            for (int issuers : {1, 2, 4})
                for (int n : {64, 96, 128}) {
                    if (mode == 1 && n == 128) continue;        // (TS leaves 448 accumulator columns: 4 x 96)
                    if (mode == 0 && n == 96) continue;
                  for (int ldtm = 0; ldtm < 2; ++ldtm) {
                    if (ldtm && (distinct == 0 || issuers != 4)) continue;
                    k_mma_multi<<<148, 256, 20 * 8192 + 1024>>>(out, n, iters, issuers, mode, distinct, ldtm);
                    cudaError_t e = cudaDeviceSynchronize();
                    long long c = 0;
                    cudaMemcpy(&c, out, 8, cudaMemcpyDeviceToHost);
                    printf("%s acc=f16 N=%3d, %d issuer warp(s), %s operands%s: %7.1f cycles / MMA aggregated (%s)\n", mode ? "TS" : "SS", n,
                           issuers, distinct ? "distinct" : "shared  ", ldtm ? " + tcgen05.ld traffic" : "", (double)c / (iters * issuers), cudaGetErrorString(e));
                  }
                }
    for (int n : {64, 128}) {
        k_lat<<<1, 128, smem>>>(out, n, 64);
        cudaError_t e = cudaDeviceSynchronize();
        long long c[2] = {0, 0};
        cudaMemcpy(c, out, 16, cudaMemcpyDeviceToHost);
        printf("stage N=%3d: 2 x MMA issue -> commit -> mbarrier seen %lld cycles;  tcgen05.ld.pack::16b (%d x x32) issue -> wait::ld %lld cycles (%s)\n",
               n, c[0], n / 64, c[1], cudaGetErrorString(e));
    }
    return 0;
}
