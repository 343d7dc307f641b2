// Signal preprocessing on the GPU: radian/preprocess.py:23-49 (mad_normalise) and :4-21
// (get_windows), call sites radian/basecall.py:78,83.
//
// One CTA per read.  Both medians are exact order statistics of small integers: the samples are
// int16, and with median = (v1 + v2) / 2 the distances |x - median| are u / 2 for the integers
// u = |2x - (v1 + v2)| <= 131070.  Each is found by a two-level counting select (a 256-bin
// histogram of the high bits, then a histogram of the low bits inside the bin that holds the
// wanted rank), warp-private histograms with match-aggregated atomics because raw signals pile up
// in two or three coarse bins.  The read is scanned five times; after the first pass it sits in L2.
// The z-scores are float64 with the reference's operation order:
// (x - median) / (1.4826 * mad), IEEE division.
#include <vector>

#include "internal.h"

namespace radian {

constexpr int kPreThreads = 256;
constexpr int kPreWarps = kPreThreads / 32;
constexpr int kDirectRange = 4096;                   // value span handled by one exact histogram
constexpr int kHistWords = 2 * kDirectRange + 64;    // >= 2 * span + 1 and >= kPreWarps * 512

// hist[w][bin] += 1 with one atomic per distinct bin of the warp
__device__ __forceinline__ void hist_add(unsigned *hist, int bin, bool valid)
{
    const unsigned act = __ballot_sync(0xffffffffu, valid);
    if (valid) {
        const unsigned peers = __match_any_sync(act, bin);
        if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[bin], __popc(peers));
    }
}

// sums the warp-private copies into copy 0 and finds, for the ranks k1 <= k2, the bin holding each
// and the rank inside that bin
template <int BINS>
__device__ void select_bins(unsigned *hist /*[kPreWarps][BINS]*/, unsigned k1, unsigned k2, int *sel /*4 ints*/)
{
    __syncthreads();
    for (int b = threadIdx.x; b < BINS; b += kPreThreads) {
        unsigned s = 0;
#pragma unroll
        for (int w = 0; w < kPreWarps; ++w) s += hist[w * BINS + b];
        hist[b] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned cum = 0;
        int b1 = -1, b2 = -1;
        unsigned r1 = 0, r2 = 0;
        for (int b = 0; b < BINS; ++b) {
            const unsigned c = hist[b];
            if (b1 < 0 && cum + c > k1) b1 = b, r1 = k1 - cum;
            if (b2 < 0 && cum + c > k2) b2 = b, r2 = k2 - cum;
            cum += c;
        }
        sel[0] = b1, sel[1] = (int)r1, sel[2] = b2, sel[3] = (int)r2;
    }
    __syncthreads();
}

// exact values at ranks k1 <= k2 of a histogram of `bins` (<= kHistWords) consecutive integers:
// every thread sums a contiguous slice, the slices are scanned, the owner of a rank walks its slice
__device__ void select_direct(const unsigned *hist, int bins, unsigned k1, unsigned k2, int *red, int *sel)
{
    __syncthreads();
    const int per = (bins + kPreThreads - 1) / kPreThreads;
    const int b0 = threadIdx.x * per;
    unsigned mine = 0;
    for (int b = b0; b < b0 + per && b < bins; ++b) mine += hist[b];
    // exclusive scan of the 256 slice sums: within the warp, then across the 8 warps
    unsigned inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned y = __shfl_up_sync(0xffffffffu, inc, o);
        if ((threadIdx.x & 31) >= o) inc += y;
    }
    if ((threadIdx.x & 31) == 31) red[threadIdx.x >> 5] = (int)inc;
    __syncthreads();
    unsigned base = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) base += (unsigned)red[w];
    const unsigned before = base + inc - mine;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
        const unsigned k = which ? k2 : k1;
        if (k >= before && k < before + mine) {
            unsigned cum = before;
            for (int b = b0; b < b0 + per && b < bins; ++b) {
                cum += hist[b];
                if (cum > k) {
                    sel[which] = b;
                    break;
                }
            }
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void hist_clear(unsigned *hist, int n)
{
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += kPreThreads) hist[i] = 0;
    __syncthreads();
}

__global__ void __launch_bounds__(kPreThreads)
normalise_kernel(const int16_t *__restrict__ signal, const int64_t *__restrict__ offsets, int n_reads,
                 double outlier, int outlier_is_int, double *__restrict__ out, int32_t *__restrict__ out_is_int,
                 int32_t *__restrict__ status)
{
    __shared__ unsigned hist[kHistWords];
    __shared__ int sel[4];
    __shared__ int pick[2];
    __shared__ int red[2 * kPreWarps];
    const int warp = threadIdx.x >> 5;
    for (int r = blockIdx.x; r < n_reads; r += gridDim.x) {
        const int64_t o0 = offsets[r];
        const int64_t n = offsets[r + 1] - o0;
        const int16_t *x = signal + o0;
        if (n == 0) {  // "Signal must not be empty to normalise" (preprocess.py:24-25)
            if (threadIdx.x == 0) status[r] = RADIAN_READ_EMPTY_SIGNAL, out_is_int[r] = 0;
            continue;
        }
        const unsigned k1 = (unsigned)((n - 1) / 2), k2 = (unsigned)(n / 2);
        const int64_t n_pad = (n + kPreThreads - 1) / kPreThreads * kPreThreads;  // whole warps for the votes

        // ---- value range: raw signals span a few hundred ADC levels, so one exact histogram of
        // the values (and one of the doubled distances) usually fits shared memory
        int lo = 32767, hi = -32768;
        for (int64_t i = threadIdx.x; i < n; i += kPreThreads) {
            const int v = x[i];
            lo = v < lo ? v : lo;
            hi = v > hi ? v : hi;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[warp] = lo, red[kPreWarps + warp] = hi;
        __syncthreads();
#pragma unroll
        for (int w = 0; w < kPreWarps; ++w) lo = min(lo, red[w]), hi = max(hi, red[kPreWarps + w]);
        if (hi - lo < kDirectRange) {
            // median of the samples
            const int nb1 = hi - lo + 1;
            hist_clear(hist, nb1);
            for (int64_t i = threadIdx.x; i < n; i += kPreThreads) atomicAdd(&hist[(int)x[i] - lo], 1u);
            select_direct(hist, nb1, k1, k2, red, sel);
            const int sum2d = sel[0] + sel[1] + 2 * lo;
            // median of the doubled distances u = |2x - (v1 + v2)| <= 2 * (hi - lo)
            const int nb2 = 2 * (hi - lo) + 1;
            hist_clear(hist, nb2);
            for (int64_t i = threadIdx.x; i < n; i += kPreThreads) atomicAdd(&hist[abs(2 * (int)x[i] - sum2d)], 1u);
            select_direct(hist, nb2, k1, k2, red, sel);
            if (threadIdx.x == 0) pick[0] = sum2d, pick[1] = sel[0] + sel[1];
            __syncthreads();
        } else {
        // ---- median of the samples: high byte, then low byte
        hist_clear(hist, kPreWarps * 256);
        for (int64_t i = threadIdx.x; i < n_pad; i += kPreThreads) {
            const bool v = i < n;
            const int key = v ? (int)x[i] + 32768 : 0;
            hist_add(hist + warp * 256, key >> 8, v);
        }
        select_bins<256>(hist, k1, k2, sel);
        const int c1 = sel[0], r1 = sel[1], c2 = sel[2], r2 = sel[3];
        for (int pass = 0; pass < (c1 == c2 ? 1 : 2); ++pass) {
            const int cb = pass == 0 ? c1 : c2;
            hist_clear(hist, kPreWarps * 256);
            for (int64_t i = threadIdx.x; i < n_pad; i += kPreThreads) {
                const int key = i < n ? (int)x[i] + 32768 : 0;
                const bool v = i < n && (key >> 8) == cb;
                hist_add(hist + warp * 256, key & 255, v);
            }
            // ranks inside the coarse bin
            const unsigned q1 = pass == 0 ? (unsigned)r1 : (unsigned)r2;
            const unsigned q2 = (c1 == c2) ? (unsigned)r2 : q1;
            select_bins<256>(hist, q1, q2, sel);
            if (threadIdx.x == 0) {
                if (pass == 0) pick[0] = (cb << 8 | sel[0]) - 32768;
                if (pass == 1 || c1 == c2) pick[1] = (cb << 8 | (c1 == c2 ? sel[2] : sel[0])) - 32768;
            }
            __syncthreads();
        }
        const int sum2 = pick[0] + pick[1];  // median = sum2 / 2 (np.median: mean of the two middle values)
        __syncthreads();

        // ---- median of the distances u / 2, u = |2x - sum2| < 2^17: high 8 bits, then low 9 bits
        hist_clear(hist, kPreWarps * 256);
        for (int64_t i = threadIdx.x; i < n_pad; i += kPreThreads) {
            const bool v = i < n;
            const int u = v ? abs(2 * (int)x[i] - sum2) : 0;
            hist_add(hist + warp * 256, u >> 9, v);
        }
        select_bins<256>(hist, k1, k2, sel);
        const int d1 = sel[0], s1 = sel[1], d2 = sel[2], s2 = sel[3];
        for (int pass = 0; pass < (d1 == d2 ? 1 : 2); ++pass) {
            const int cb = pass == 0 ? d1 : d2;
            hist_clear(hist, kPreWarps * 512);
            for (int64_t i = threadIdx.x; i < n_pad; i += kPreThreads) {
                const int u = i < n ? abs(2 * (int)x[i] - sum2) : 0;
                const bool v = i < n && (u >> 9) == cb;
                hist_add(hist + warp * 512, u & 511, v);
            }
            const unsigned q1 = pass == 0 ? (unsigned)s1 : (unsigned)s2;
            const unsigned q2 = (d1 == d2) ? (unsigned)s2 : q1;
            select_bins<512>(hist, q1, q2, sel);
            if (threadIdx.x == 0) {
                if (pass == 0) pick[0] = cb << 9 | sel[0];
                if (pass == 1 || d1 == d2) pick[1] = cb << 9 | (d1 == d2 ? sel[2] : sel[0]);
            }
            __syncthreads();
        }
        const int us = pick[0] + pick[1];
        __syncthreads();
        if (threadIdx.x == 0) pick[0] = sum2, pick[1] = us;
        __syncthreads();
        }
        const int sum2 = pick[0];   // median = sum2 / 2 (np.median: mean of the two middle values)
        const int usum = pick[1];   // mad = (u1/2 + u2/2) / 2
        __syncthreads();
        if (usum == 0) {  // "MAD is zero, issue with signal." (preprocess.py:47-48)
            if (threadIdx.x == 0) status[r] = RADIAN_READ_MAD_ZERO, out_is_int[r] = 0;
            continue;
        }
        const double median = (double)sum2 / 2.0;
        const double mad = (double)usum / 4.0;
        const double scale = __dmul_rn(1.4826, mad);
        // np.vectorize takes the output dtype from the first element: a clipped first sample with an
        // integer outlier_z_score makes the whole result int64 (values truncated towards zero)
        const double z0 = __ddiv_rn(__dsub_rn((double)x[0], median), scale);
        const bool as_int = outlier_is_int && (z0 > outlier || z0 < -outlier);
        double *o = out + o0;
        for (int64_t i = threadIdx.x; i < n; i += kPreThreads) {
            double z = __ddiv_rn(__dsub_rn((double)x[i], median), scale);
            if (z > outlier) z = outlier;
            else if (z < -outlier) z = -outlier;
            if (as_int)
                reinterpret_cast<long long *>(o)[i] = (long long)z;
            else
                o[i] = z;
        }
        if (threadIdx.x == 0) status[r] = RADIAN_READ_OK, out_is_int[r] = as_int ? 1 : 0;
    }
}

// get_windows: window w of a read starts at sample w * step; the last one is zero padded
__global__ void windows_kernel(const double *__restrict__ norm, const int64_t *__restrict__ offsets,
                               const int64_t *__restrict__ win_off, int n_reads, int window, int step,
                               double *__restrict__ out)
{
    for (int r = blockIdx.y; r < n_reads; r += gridDim.y) {
        const int64_t n = offsets[r + 1] - offsets[r];
        const double *x = norm + offsets[r];
        const int64_t total = (win_off[r + 1] - win_off[r]) * window;
        double *o = out + win_off[r] * window;
        for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
             e += (int64_t)gridDim.x * blockDim.x) {
            const int64_t w = e / window, i = e - w * window;
            const int64_t t = w * step + i;
            o[e] = t < n ? x[t] : 0.0;
        }
    }
}

}  // namespace radian

using namespace radian;

extern "C" int radian_windows_plan(const int64_t *offsets, int n_reads, int window, int step, int64_t *n_windows,
                                   int32_t *pad_end)
{
    if (!offsets || !n_windows || !pad_end || n_reads < 0) {
        set_error("radian_windows_plan: null argument");
        return RADIAN_E_ARG;
    }
    if (step <= 0) {
        set_error("Step size must be > 0");
        return RADIAN_E_ARG;
    }
    if (step > window) {
        set_error("Step size must be <= window size");
        return RADIAN_E_ARG;
    }
    for (int r = 0; r < n_reads; ++r) {
        const int64_t n = offsets[r + 1] - offsets[r];
        if (n < 0) {
            set_error("radian_windows_plan: offsets not monotone at read %d", r);
            return RADIAN_E_ARG;
        }
        // full windows: starts 0, step, ... while start + window <= n (preprocess.py:11-14)
        const int64_t full = n >= window ? (n - window) / step + 1 : 0;
        const int64_t start = full * step;
        n_windows[r] = full + 1;
        pad_end[r] = (int32_t)(window - (n - start));
    }
    return RADIAN_OK;
}

extern "C" int radian_normalise_batch_dev(const int16_t *signal, const int64_t *offsets, int n_reads,
                                          double outlier_z_score, int outlier_is_int, void *out,
                                          int32_t *out_is_int64, int32_t *out_status, radian_stream_t stream)
{
    if (n_reads < 0 || !offsets || !out_is_int64 || !out_status) {
        set_error("radian_normalise_batch_dev: null argument");
        return RADIAN_E_ARG;
    }
    if (n_reads == 0) return RADIAN_OK;
    int device = 0;
    RADIAN_CUDA(cudaGetDevice(&device));
    DeviceInfo di;
    int rc = device_info(device, &di);
    if (rc) return rc;
    const int grid = n_reads < di.sm_count * 8 ? n_reads : di.sm_count * 8;
    normalise_kernel<<<grid, kPreThreads, 0, (cudaStream_t)stream>>>(signal, offsets, n_reads, outlier_z_score,
                                                                       outlier_is_int, (double *)out, out_is_int64,
                                                                       out_status);
    RADIAN_CUDA(cudaGetLastError());
    return RADIAN_OK;
}

extern "C" int radian_windows_batch_dev(const double *norm, const int64_t *offsets, const int64_t *window_offsets,
                                        int n_reads, int window, int step, double *out, radian_stream_t stream)
{
    if (n_reads < 0 || !offsets || !window_offsets || (n_reads > 0 && !out)) {
        set_error("radian_windows_batch_dev: null argument");
        return RADIAN_E_ARG;
    }
    if (step <= 0 || step > window) {
        set_error(step <= 0 ? "Step size must be > 0" : "Step size must be <= window size");
        return RADIAN_E_ARG;
    }
    if (n_reads == 0) return RADIAN_OK;
    const dim3 grid(32, (unsigned)(n_reads < 4096 ? n_reads : 4096));
    windows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(norm, offsets, window_offsets, n_reads, window, step, out);
    RADIAN_CUDA(cudaGetLastError());
    return RADIAN_OK;
}

extern "C" int radian_normalise_batch_host(const int16_t *signal, const int64_t *offsets, int n_reads,
                                           double outlier_z_score, int outlier_is_int, void *out,
                                           int32_t *out_is_int64, int32_t *out_status, int device)
{
    if (n_reads < 0 || !offsets || !out_is_int64 || !out_status) {
        set_error("radian_normalise_batch_host: null argument");
        return RADIAN_E_ARG;
    }
    if (n_reads == 0) return RADIAN_OK;
    if (radian_device_count() <= device || device < 0) {
        set_error("radian_normalise_batch_host: CUDA device %d not available (no CPU fallback exists)", device);
        return RADIAN_E_CUDA;
    }
    for (int r = 0; r < n_reads; ++r)
        if (offsets[r + 1] < offsets[r]) {
            set_error("radian_normalise_batch_host: offsets not monotone at read %d", r);
            return RADIAN_E_ARG;
        }
    const int64_t total = offsets[n_reads] - offsets[0];
    std::lock_guard<std::mutex> host_lock(host_mutex(device));
    RADIAN_CUDA(cudaSetDevice(device));
    cudaMemPool_t pool_ = nullptr;  // this library's own stream-ordered pool on the device
    {
        int krc = keep_pool(device, &pool_);
        if (krc) return krc;
    }
    cudaStream_t st = nullptr;
    RADIAN_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    int16_t *d_sig = nullptr;
    int64_t *d_off = nullptr;
    double *d_out = nullptr;
    int32_t *d_int = nullptr, *d_status = nullptr;
    std::vector<int64_t> rel((size_t)n_reads + 1);
    for (int r = 0; r <= n_reads; ++r) rel[r] = offsets[r] - offsets[0];
    int ret = RADIAN_OK;
    cudaError_t e;
#define TRY(x)                                   \
    if (ret == RADIAN_OK && (e = (x)) != cudaSuccess) ret = cuda_fail(e, #x)
    TRY(cudaMallocFromPoolAsync(&d_sig, (size_t)(total ? total : 1) * 2, pool_, st));
    TRY(cudaMallocFromPoolAsync(&d_off, (size_t)(n_reads + 1) * 8, pool_, st));
    TRY(cudaMallocFromPoolAsync(&d_out, (size_t)(total ? total : 1) * 8, pool_, st));
    TRY(cudaMallocFromPoolAsync(&d_int, (size_t)n_reads * 4, pool_, st));
    TRY(cudaMallocFromPoolAsync(&d_status, (size_t)n_reads * 4, pool_, st));
    if (total) TRY(cudaMemcpyAsync(d_sig, signal + offsets[0], (size_t)total * 2, cudaMemcpyHostToDevice, st));
    TRY(cudaMemcpyAsync(d_off, rel.data(), (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    if (ret == RADIAN_OK)
        ret = radian_normalise_batch_dev(d_sig, d_off, n_reads, outlier_z_score, outlier_is_int, d_out, d_int,
                                         d_status, st);
    if (total) TRY(cudaMemcpyAsync((double *)out + offsets[0], d_out, (size_t)total * 8, cudaMemcpyDeviceToHost, st));
    TRY(cudaMemcpyAsync(out_is_int64, d_int, (size_t)n_reads * 4, cudaMemcpyDeviceToHost, st));
    TRY(cudaMemcpyAsync(out_status, d_status, (size_t)n_reads * 4, cudaMemcpyDeviceToHost, st));
    TRY(cudaStreamSynchronize(st));
#undef TRY
    void *frees[] = {d_sig, d_off, d_out, d_int, d_status};
    for (void *p : frees)
        if (p) cudaFreeAsync(p, st);
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    return ret;
}

extern "C" int radian_windows_batch_host(const double *norm, const int64_t *offsets, int n_reads, int window,
                                         int step, const int64_t *window_offsets, double *out, int device)
{
    if (n_reads < 0 || !offsets || !window_offsets || (n_reads > 0 && !out)) {
        set_error("radian_windows_batch_host: null argument");
        return RADIAN_E_ARG;
    }
    if (step <= 0 || step > window) {
        set_error(step <= 0 ? "Step size must be > 0" : "Step size must be <= window size");
        return RADIAN_E_ARG;
    }
    if (n_reads == 0) return RADIAN_OK;
    if (radian_device_count() <= device || device < 0) {
        set_error("radian_windows_batch_host: CUDA device %d not available (no CPU fallback exists)", device);
        return RADIAN_E_CUDA;
    }
    const int64_t total = offsets[n_reads] - offsets[0];
    const int64_t n_win = window_offsets[n_reads] - window_offsets[0];
    std::lock_guard<std::mutex> host_lock(host_mutex(device));
    RADIAN_CUDA(cudaSetDevice(device));
    cudaMemPool_t pool_ = nullptr;  // this library's own stream-ordered pool on the device
    {
        int krc = keep_pool(device, &pool_);
        if (krc) return krc;
    }
    cudaStream_t st = nullptr;
    RADIAN_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    double *d_in = nullptr, *d_out = nullptr;
    int64_t *d_off = nullptr, *d_woff = nullptr;
    std::vector<int64_t> rel((size_t)n_reads + 1), wrel((size_t)n_reads + 1);
    for (int r = 0; r <= n_reads; ++r) rel[r] = offsets[r] - offsets[0], wrel[r] = window_offsets[r] - window_offsets[0];
    int ret = RADIAN_OK;
    cudaError_t e;
#define TRY(x)                                   \
    if (ret == RADIAN_OK && (e = (x)) != cudaSuccess) ret = cuda_fail(e, #x)
    TRY(cudaMallocFromPoolAsync(&d_in, (size_t)(total ? total : 1) * 8, pool_, st));
    TRY(cudaMallocFromPoolAsync(&d_out, (size_t)(n_win ? n_win : 1) * window * 8, pool_, st));
    TRY(cudaMallocFromPoolAsync(&d_off, (size_t)(n_reads + 1) * 8, pool_, st));
    TRY(cudaMallocFromPoolAsync(&d_woff, (size_t)(n_reads + 1) * 8, pool_, st));
    if (total) TRY(cudaMemcpyAsync(d_in, norm + offsets[0], (size_t)total * 8, cudaMemcpyHostToDevice, st));
    TRY(cudaMemcpyAsync(d_off, rel.data(), (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    TRY(cudaMemcpyAsync(d_woff, wrel.data(), (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    if (ret == RADIAN_OK && n_win > 0)
        ret = radian_windows_batch_dev(d_in, d_off, d_woff, n_reads, window, step, d_out, st);
    if (n_win > 0)
        TRY(cudaMemcpyAsync(out + window_offsets[0] * window, d_out, (size_t)n_win * window * 8,
                            cudaMemcpyDeviceToHost, st));
    TRY(cudaStreamSynchronize(st));
#undef TRY
    void *frees[] = {d_in, d_out, d_off, d_woff};
    for (void *p : frees)
        if (p) cudaFreeAsync(p, st);
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    return ret;
}
