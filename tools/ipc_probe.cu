// Probe: can two processes on one box map each other's cudaMalloc memory (legacy CUDA IPC over NVLink P2P) and
// exchange data + flags with plain loads/stores from kernels?  Run with >= 2 visible GPUs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ipc_probe tools/ipc_probe.cu && ./ipc_probe
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <unistd.h>
#include <sys/wait.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("[%d] %s -> %s\n", rank, #x, cudaGetErrorString(e)); return 2; } } while (0)

__global__ void k_fill(unsigned long long* p, int n, unsigned long long v) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = v + i;
}
// signal the peer, wait for the peer's signal, then sum the peer's buffer
__global__ void k_exchange(unsigned* my_flag, unsigned* peer_flag, const unsigned long long* peer, int n, unsigned epoch,
                           unsigned long long* out, long long* cycles) {
    __shared__ int ok;
    if (threadIdx.x == 0) {
        long long t0 = clock64();
        if (blockIdx.x == 0) {
            __threadfence_system();
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_flag), "r"(epoch) : "memory");
        }
        unsigned v = 0;
        ok = 1;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(my_flag) : "memory");
            if (clock64() - t0 > 4000000000ll) { ok = 0; break; }
        } while (v < epoch);
        if (blockIdx.x == 0) cycles[0] = clock64() - t0;
    }
    __syncthreads();
    if (!ok) return;
    unsigned long long s = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        unsigned long long v;
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(peer + i) : "memory");
        s += v;
    }
    atomicAdd(out, s);
}

int main() {
    int p2c[2], c2p[2];
    if (pipe(p2c) || pipe(c2p)) return 1;
    pid_t pid = fork();
    const int rank = pid == 0 ? 1 : 0;
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev < 2) { if (rank == 0) printf("need 2 GPUs, have %d\n", ndev); return 0; }
    CK(cudaSetDevice(rank));
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, rank, 1 - rank));
    printf("[%d] canAccessPeer=%d\n", rank, can);
    const int n = 1 << 18;   // 2 MiB of int64
    unsigned long long* buf; unsigned* flags; unsigned long long* out; long long* cyc;
    CK(cudaMalloc(&buf, sizeof(unsigned long long) * n + 4096));
    CK(cudaMemset(buf, 0, sizeof(unsigned long long) * n + 4096));
    flags = reinterpret_cast<unsigned*>(buf + n);
    CK(cudaMalloc(&out, 8)); CK(cudaMalloc(&cyc, 8));
    cudaIpcMemHandle_t mine, theirs;
    CK(cudaIpcGetMemHandle(&mine, buf));
    const int wfd = rank == 0 ? p2c[1] : c2p[1], rfd = rank == 0 ? c2p[0] : p2c[0];
    if (write(wfd, &mine, sizeof(mine)) != sizeof(mine)) return 3;
    if (read(rfd, &theirs, sizeof(theirs)) != sizeof(theirs)) return 3;
    void* peer_v = nullptr;
    CK(cudaIpcOpenMemHandle(&peer_v, theirs, cudaIpcMemLazyEnablePeerAccess));
    unsigned long long* peer = static_cast<unsigned long long*>(peer_v);
    unsigned* peer_flags = reinterpret_cast<unsigned*>(peer + n);
    printf("[%d] opened peer buffer %p\n", rank, peer_v);
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (unsigned epoch = 1; epoch <= 6; ++epoch) {
        k_fill<<<148, 256>>>(buf, n, 1000ull * rank + epoch);
        CK(cudaMemset(out, 0, 8));
        CK(cudaEventRecord(a));
        k_exchange<<<148, 512>>>(flags, peer_flags, peer, n, epoch, out, cyc);
        CK(cudaEventRecord(b));
        CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        unsigned long long got; long long c;
        CK(cudaMemcpy(&got, out, 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
        const unsigned long long want = (unsigned long long)n * (1000ull * (1 - rank) + epoch) + (unsigned long long)n * (n - 1) / 2;
        printf("[%d] epoch %u: %s  kernel %.1f us (flag wait %lld cycles)  pulled %.1f MB\n", rank, epoch,
               got == want ? "OK" : "MISMATCH", ms * 1e3f, c, n * 8 / 1e6);
        // both sides must finish reading before the next fill: handshake through the pipe
        char t = 1;
        if (write(wfd, &t, 1) != 1 || read(rfd, &t, 1) != 1) return 4;
    }
    CK(cudaIpcCloseMemHandle(peer_v));
    if (rank == 0) { int st; waitpid(pid, &st, 0); }
    return 0;
}
