// Probe: what limits a chunked pinned H2D stream with per-chunk "ready" flags while a persistent
// kernel polls the flag?  (design input for radian_decode_batch_host; not part of the library)
#include <cuda.h>
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstring>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s line %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(1);} } while (0)
__global__ void poller(const volatile int *flag, int target, unsigned long long *sink)
{
    unsigned long long n = 0;
    while (*flag < target) { __nanosleep(400); ++n; }
    if (threadIdx.x == 0 && blockIdx.x == 0) *sink = n;
}
typedef CUresult (*WriteValFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char **argv)
{
    size_t total = (size_t)4 << 30;
    char *h, *d; int *dflag; int *hflag; unsigned long long *sink;
    CK(cudaHostAlloc(&h, total, cudaHostAllocDefault));
    CK(cudaMalloc(&d, total));
    CK(cudaMalloc(&dflag, 256)); CK(cudaMalloc(&sink, 8));
    CK(cudaHostAlloc(&hflag, 1 << 22, cudaHostAllocDefault));
    for (int i = 0; i < (1 << 20); ++i) hflag[i] = i + 1;
    memset(h, 1, total);
    cudaStream_t cs, ks; CK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&ks, cudaStreamNonBlocking));
    WriteValFn wv = nullptr; cudaDriverEntryPointQueryResult qr;
    CK(cudaGetDriverEntryPoint("cuStreamWriteValue32", (void **)&wv, cudaEnableDefault, &qr));
    printf("cuStreamWriteValue32 %s\n", wv ? "found" : "missing");
    for (int rep = 0; rep < 2; ++rep) {
        double t0 = now(); CK(cudaMemcpyAsync(d, h, total, cudaMemcpyHostToDevice, cs)); CK(cudaStreamSynchronize(cs));
        double dt = now() - t0; printf("one copy: %.1f GB/s\n", total / dt / 1e9);
    }
    size_t sizes[] = {(size_t)128 << 10, (size_t)1 << 20, (size_t)8 << 20};
    for (size_t cs_bytes : sizes) {
        int n = (int)(total / cs_bytes);
        for (int mode = 0; mode < 6; ++mode) {
            // 0 chunks only, 1 chunks+flag memcpy, 2 chunks+writevalue, 3/4/5 = same with poller kernel running
            const bool poll = mode >= 3; const int fm = mode % 3;
            CK(cudaMemset(dflag, 0, 4)); CK(cudaDeviceSynchronize());
            if (poll) { poller<<<148 * 6, 128, 0, ks>>>(dflag, fm == 0 ? 1 : n, sink); }
            double t0 = now();
            for (int i = 0; i < n; ++i) {
                CK(cudaMemcpyAsync(d + (size_t)i * cs_bytes, h + (size_t)i * cs_bytes, cs_bytes, cudaMemcpyHostToDevice, cs));
                if (fm == 1) CK(cudaMemcpyAsync(dflag, &hflag[i], 4, cudaMemcpyHostToDevice, cs));
                if (fm == 2) { if (wv((CUstream)cs, (CUdeviceptr)dflag, (cuuint32_t)(i + 1), 0) != CUDA_SUCCESS) { printf("wv failed\n"); return 1; } }
            }
            double tsub = now() - t0;
            if (fm == 0) CK(cudaMemcpyAsync(dflag, &hflag[0], 4, cudaMemcpyHostToDevice, cs));
            CK(cudaStreamSynchronize(cs));
            double dt = now() - t0;
            CK(cudaDeviceSynchronize());
            printf("chunk %7zu KB x %6d mode %d (%s%s): %.1f GB/s (submit %.3fs, total %.3fs)\n", cs_bytes >> 10, n, mode,
                   fm == 0 ? "copies" : fm == 1 ? "copies+flagcpy" : "copies+writeval", poll ? "+poller" : "", total / dt / 1e9, tsub, dt);
        }
    }
    return 0;
}
