// Latency / throughput probe for the instructions on the decode kernel's per-frame dependency chain
// (sm_100a): nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat_probe lat_probe.cu && ./lat_probe
// Prints cycles per dependent operation for one warp, and cycles per warp-instruction when 1..8 warps
// of one scheduler issue independent operations (pipe throughput).
#include <cuda_runtime.h>
#include <stdio.h>

constexpr int N = 4096;

__global__ void k_dep(double *out, long long *cyc, double seed)
{
    __shared__ double sm[64];
    __shared__ unsigned smi[64];
    const int lane = threadIdx.x & 31;
    sm[lane] = seed + lane;
    sm[lane + 32] = 0.0;
    smi[lane] = lane;
    __syncwarp();
    double x = seed + lane * 1e-3, y = 1.0000001;
    long long t0, t1;
    // DMUL chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = __dmul_rn(x, y);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    // DADD chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = __dadd_rn(x, y);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[1] = t1 - t0;
    // SHFL chain (32-bit)
    unsigned u = __double2hiint(x) + lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) u = __shfl_sync(0xffffffffu, u, (u + 1) & 31);
    t1 = clock64();
    if (threadIdx.x == 0) cyc[2] = t1 - t0;
    // LDS chain (pointer chasing through shared memory)
    unsigned idx = lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) idx = smi[idx & 31];
    t1 = clock64();
    if (threadIdx.x == 0) cyc[3] = t1 - t0;
    // STS -> syncwarp -> LDS round trip (what the copy/extend merge does every frame)
    double z = x;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) {
        sm[lane] = z;
        __syncwarp();
        z = sm[(lane + 1) & 31] + 1.0;
        __syncwarp();
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[4] = t1 - t0;
    // VOTE.ALL + branch chain
    int acc = 0;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; ++i) {
        if (!__all_sync(0xffffffffu, (u + i) != 0x12345678u)) break;
        acc += i;
    }
    t1 = clock64();
    if (threadIdx.x == 0) cyc[5] = t1 - t0;
    // integer compare/select chain (IMAD/ISETP class)
    int q = lane;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) q = max(q + 3, (int)u) & 0x7fffffff;
    t1 = clock64();
    if (threadIdx.x == 0) cyc[6] = t1 - t0;
    out[threadIdx.x] = x + z + u + idx + acc + q;
}

// independent DMULs from `warps` warps on one SM: cycles per warp-instruction per scheduler
__global__ void k_tp(double *out, long long *cyc, double seed)
{
    double a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = seed + threadIdx.x + j;
    const double y = 1.0000001;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = __dmul_rn(a[j], y);
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += a[j];
    out[threadIdx.x] = s;
}

int main()
{
    double *out;
    long long *cyc, h[8];
    cudaMalloc(&out, 1024 * 8);
    cudaMalloc(&cyc, 64);
    k_dep<<<1, 32>>>(out, cyc, 1.5);
    k_dep<<<1, 32>>>(out, cyc, 1.5);
    cudaMemcpy(h, cyc, 56, cudaMemcpyDeviceToHost);
    const char *names[] = {"DMUL dependent", "DADD dependent", "SHFL.IDX dependent", "LDS dependent (pointer chase)",
                           "STS+syncwarp+LDS+DADD+syncwarp", "VOTE.ALL + branch + IADD", "IADD+VIMNMX+LOP dependent"};
    for (int i = 0; i < 7; ++i) printf("%-36s %.1f cycles/iteration\n", names[i], (double)h[i] / N);
    for (int warps = 4; warps <= 32; warps *= 2) {
        k_tp<<<1, warps * 32>>>(out, cyc, 1.5);
        cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost);
        // warps/4 warps per scheduler, each issuing N*8 DMULs
        printf("DMUL throughput, %2d warps on the SM: %.2f cycles per warp-DMUL per scheduler\n", warps,
               (double)h[0] / ((double)N * 8 * (warps / 4)));
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
