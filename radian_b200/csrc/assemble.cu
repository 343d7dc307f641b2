// Merge of overlapping chunk posteriors, sm_100a.
//
// Reference: radian/matrix_assembly.py:6-53.  create_vstack (:12-34) stacks chunk k at global
// row k*step; collapse_vstack (:36-44) keeps one row per global row: the row of the FIRST chunk
// covering it, because average_dist (:46-53) discards the result of np.add (:52).  Rows covered
// by more than one chunk go through sklearn normalize([row], norm="l1") (:53): cast to float64,
// divided by sum(|x|) accumulated left to right, a norm below 10*eps replaced by 1.  Rows
// covered once are passed through untouched.
//
// Pure streaming: 20 B read + 20/40 B written per global row, no reuse -> HBM bound.  One thread
// owns one output row; a warp therefore reads a contiguous 640 B span of the winning chunk and
// writes a contiguous 640/1280 B span of the output.
#include "internal.h"

namespace radian {

struct AssembleArgs {
    const float *chunks;
    const int64_t *chunk_row_offsets;
    const int64_t *read_chunk_ranges;
    const int64_t *out_row_offsets;
    int n_reads;
    int step;
    int max_chunk_rows;
    void *out;
    int64_t total_rows;
};

template <bool F64>
__global__ void __launch_bounds__(256) assemble_kernel(const AssembleArgs a)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < a.total_rows; g += stride) {
        // read owning global output row g: binary search in out_row_offsets
        int lo = 0, hi = a.n_reads;
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (__ldg(a.out_row_offsets + mid) <= g) lo = mid; else hi = mid;
        }
        const int r = lo;
        const int64_t t = g - __ldg(a.out_row_offsets + r);
        const int64_t c0 = __ldg(a.read_chunk_ranges + r);
        const int64_t nchunk = __ldg(a.read_chunk_ranges + r + 1) - c0;
        // chunks that can cover row t: k*step <= t < k*step + rows(k), rows(k) <= max_chunk_rows
        int64_t k_hi = t / a.step;
        if (k_hi > nchunk - 1) k_hi = nchunk - 1;
        int64_t k_lo = (t - a.max_chunk_rows + a.step) / a.step;  // ceil((t-max+1)/step)
        if (t - a.max_chunk_rows + 1 <= 0) k_lo = 0;
        int64_t src_row = -1;
        int cover = 0;
        for (int64_t k = k_lo; k <= k_hi && cover < 2; ++k) {
            const int64_t ro = __ldg(a.chunk_row_offsets + c0 + k);
            const int64_t rows = __ldg(a.chunk_row_offsets + c0 + k + 1) - ro;
            const int64_t j = t - k * a.step;
            if (j >= 0 && j < rows) {
                if (cover == 0) src_row = ro + j;
                ++cover;
            }
        }
        const float *src = a.chunks + src_row * 5;
        float x[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) x[i] = __ldcs(src + i);
        if (F64) {
            double *dst = (double *)a.out + g * 5;
            double y[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) y[i] = (double)x[i];
            if (cover > 1) {
                double nrm = 0.0;
#pragma unroll
                for (int i = 0; i < 5; ++i) nrm = __dadd_rn(nrm, fabs(y[i]));
                if (nrm < 10.0 * 2.220446049250313e-16) nrm = 1.0;
#pragma unroll
                for (int i = 0; i < 5; ++i) y[i] = y[i] / nrm;
            }
#pragma unroll
            for (int i = 0; i < 5; ++i) __stcs(dst + i, y[i]);
        } else {
            float *dst = (float *)a.out + g * 5;
#pragma unroll
            for (int i = 0; i < 5; ++i) __stcs(dst + i, x[i]);
        }
    }
}

int assemble_launch(const AssembleArgs &a, bool f64, int device, cudaStream_t stream)
{
    if (a.total_rows == 0) return 0;
    DeviceInfo di;
    int rc = device_info(device, &di);
    if (rc) return rc;
    const int block = 256;
    int64_t need = (a.total_rows + block - 1) / block;
    int64_t maxg = (int64_t)di.sm_count * 8;  // 8 resident CTAs of 256 threads per SM
    int grid = (int)(need < maxg ? need : maxg);
    if (f64)
        assemble_kernel<true><<<grid, block, 0, stream>>>(a);
    else
        assemble_kernel<false><<<grid, block, 0, stream>>>(a);
    RADIAN_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace radian

using namespace radian;

extern "C" int radian_assemble_plan(const int64_t *chunk_row_offsets, const int64_t *read_chunk_ranges,
                                    int n_reads, int step, int64_t *out_rows_per_read,
                                    int *out_any_overlap, int32_t *out_max_chunk_rows)
{
    if (step <= 0 || n_reads < 0 || !chunk_row_offsets || !read_chunk_ranges) {
        set_error("radian_assemble_plan: bad arguments (step must be > 0)");
        return RADIAN_E_ARG;
    }
    int any = 0;
    int64_t maxrows = 0;
    for (int r = 0; r < n_reads; ++r) {
        int64_t T = 0;
        const int64_t c0 = read_chunk_ranges[r], c1 = read_chunk_ranges[r + 1];
        for (int64_t c = c0; c < c1; ++c) {
            const int64_t rows = chunk_row_offsets[c + 1] - chunk_row_offsets[c];
            if (rows < 0) {
                set_error("radian_assemble_plan: chunk_row_offsets not monotone at chunk %lld", (long long)c);
                return RADIAN_E_ARG;
            }
            if (rows == 0) continue;
            const int64_t start = (c - c0) * (int64_t)step;
            if (start > T) {  // create_vstack appends one row at a time: IndexError in the reference
                set_error("radian_assemble_plan: read %d chunk %lld starts at row %lld but only %lld rows exist",
                          r, (long long)(c - c0), (long long)start, (long long)T);
                return RADIAN_E_GAP;
            }
            if (start < T) any = 1;
            if (start + rows > T) T = start + rows;
            if (rows > maxrows) maxrows = rows;
        }
        if (out_rows_per_read) out_rows_per_read[r] = T;
    }
    if (out_any_overlap) *out_any_overlap = any;
    if (out_max_chunk_rows) *out_max_chunk_rows = (int32_t)maxrows;
    return RADIAN_OK;
}

extern "C" int radian_assemble_batch_dev(const float *chunks, const int64_t *chunk_row_offsets,
                                         const int64_t *read_chunk_ranges, const int64_t *out_row_offsets,
                                         int n_reads, int step, int32_t max_chunk_rows,
                                         int64_t total_out_rows, void *out, int out_is_f64,
                                         radian_stream_t stream)
{
    if (n_reads < 0 || step <= 0 || max_chunk_rows < 0 || total_out_rows < 0) {
        set_error("radian_assemble_batch_dev: bad arguments");
        return RADIAN_E_ARG;
    }
    if (n_reads == 0) return RADIAN_OK;
    int device = 0;
    RADIAN_CUDA(cudaGetDevice(&device));
    AssembleArgs a{chunks, chunk_row_offsets, read_chunk_ranges, out_row_offsets, n_reads, step,
                   max_chunk_rows < 1 ? 1 : max_chunk_rows, out, total_out_rows};
    return assemble_launch(a, out_is_f64 != 0, device, (cudaStream_t)stream);
}

extern "C" int radian_assemble_batch_host(const float *chunks, const int64_t *chunk_row_offsets,
                                          const int64_t *read_chunk_ranges, const int64_t *out_row_offsets,
                                          int n_reads, int step, void *out, int out_is_f64, int device)
{
    if (n_reads < 0 || step <= 0) {
        set_error("radian_assemble_batch_host: bad arguments");
        return RADIAN_E_ARG;
    }
    if (n_reads == 0) return RADIAN_OK;
    int any = 0;
    int32_t maxrows = 0;
    int rc = radian_assemble_plan(chunk_row_offsets, read_chunk_ranges, n_reads, step, nullptr, &any, &maxrows);
    if (rc) return rc;
    if (any && !out_is_f64) {
        set_error("radian_assemble_batch_host: overlapping chunks need a float64 output (matrix_assembly.py:53)");
        return RADIAN_E_ARG;
    }
    RADIAN_CUDA(cudaSetDevice(device));
    {
        int krc = keep_pool(device);
        if (krc) return krc;
    }
    const int64_t n_chunks = read_chunk_ranges[n_reads];
    const int64_t in_rows = chunk_row_offsets[n_chunks];
    const int64_t out_rows = out_row_offsets[n_reads];
    const size_t out_bytes = (size_t)out_rows * 5 * (out_is_f64 ? 8 : 4);
    cudaStream_t st;
    RADIAN_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    float *d_chunks = nullptr;
    int64_t *d_cro = nullptr, *d_rcr = nullptr, *d_oro = nullptr;
    void *d_out = nullptr;
    int ret = RADIAN_OK;
    cudaError_t e;
#define TRY(x)                                   \
    if (ret == RADIAN_OK && (e = (x)) != cudaSuccess) ret = cuda_fail(e, #x)
    TRY(cudaMallocAsync(&d_chunks, (size_t)(in_rows ? in_rows : 1) * 20, st));
    TRY(cudaMallocAsync(&d_cro, (size_t)(n_chunks + 1) * 8, st));
    TRY(cudaMallocAsync(&d_rcr, (size_t)(n_reads + 1) * 8, st));
    TRY(cudaMallocAsync(&d_oro, (size_t)(n_reads + 1) * 8, st));
    TRY(cudaMallocAsync(&d_out, out_bytes ? out_bytes : 1, st));
    TRY(cudaMemcpyAsync(d_chunks, chunks, (size_t)in_rows * 20, cudaMemcpyHostToDevice, st));
    TRY(cudaMemcpyAsync(d_cro, chunk_row_offsets, (size_t)(n_chunks + 1) * 8, cudaMemcpyHostToDevice, st));
    TRY(cudaMemcpyAsync(d_rcr, read_chunk_ranges, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    TRY(cudaMemcpyAsync(d_oro, out_row_offsets, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    if (ret == RADIAN_OK) {
        AssembleArgs a{d_chunks, d_cro, d_rcr, d_oro, n_reads, step, maxrows < 1 ? 1 : maxrows, d_out, out_rows};
        ret = assemble_launch(a, out_is_f64 != 0, device, st);
    }
    TRY(cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, st));
    TRY(cudaStreamSynchronize(st));
#undef TRY
    if (d_chunks) cudaFreeAsync(d_chunks, st);
    if (d_cro) cudaFreeAsync(d_cro, st);
    if (d_rcr) cudaFreeAsync(d_rcr, st);
    if (d_oro) cudaFreeAsync(d_oro, st);
    if (d_out) cudaFreeAsync(d_out, st);
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    return ret;
}
