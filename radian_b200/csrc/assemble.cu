// Merge of overlapping chunk posteriors, sm_100a.
//
// Reference: radian/matrix_assembly.py:6-53.  create_vstack (:12-34) stacks chunk k at global
// row k*step; collapse_vstack (:36-44) keeps one row per global row: the row of the FIRST chunk
// covering it, because average_dist (:46-53) discards the result of np.add (:52).  Rows covered
// by more than one chunk go through sklearn normalize([row], norm="l1") (:53): cast to float64,
// divided by sum(|x|) accumulated left to right, a norm below 10*eps replaced by 1.  Rows
// covered once are passed through untouched.
//
// Pure streaming: 20 B read + 20/40 B written per global row, no reuse -> HBM bound.
#include <type_traits>

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
    int uniform;  // every chunk but each read's last has max_chunk_rows rows
};

// which read owns global output row g (out_row_offsets is sorted)
__device__ __forceinline__ int find_read(const int64_t *__restrict__ oro, int n_reads, int64_t g)
{
    int lo = 0, hi = n_reads;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(oro + mid) <= g) lo = mid; else hi = mid;
    }
    return lo;
}

constexpr int kWarps = 8;       // warps per CTA, each with its own tile: no CTA barrier
constexpr int kRowsPerLane = 2; // independent rows per lane and iteration (memory-level parallelism)
constexpr int kTileRows = 32 * kRowsPerLane;

// where output row t of a read comes from: source row in `chunks` and how many chunks cover it
__device__ __forceinline__ void locate_row(const AssembleArgs &a, int r, int64_t t, int64_t &src_row, int &cover)
{
    const int64_t W = a.max_chunk_rows;
    const int64_t c0 = __ldg(a.read_chunk_ranges + r);
    const int64_t nchunk = __ldg(a.read_chunk_ranges + r + 1) - c0;
    // chunks that can cover row t: k*step <= t < k*step + rows(k), rows(k) <= W
    int64_t k_hi = t / a.step;
    if (k_hi > nchunk - 1) k_hi = nchunk - 1;
    int64_t k_lo = (t - W + a.step) / a.step;  // ceil((t-W+1)/step)
    if (t - W + 1 <= 0) k_lo = 0;
    src_row = 0;
    cover = 0;
    if (a.uniform) {
        // every chunk but the read's last has exactly W rows (the reference's windowing,
        // preprocess.py:4-22): chunks k_lo..k_hi all cover t, except possibly the last one
        const int64_t ro0 = __ldg(a.chunk_row_offsets + c0);
        int64_t k_top = k_hi;
        if (k_hi == nchunk - 1) {
            const int64_t last_rows = __ldg(a.chunk_row_offsets + c0 + nchunk) - (ro0 + (nchunk - 1) * W);
            if (t - k_hi * a.step >= last_rows) k_top = k_hi - 1;
        }
        cover = (int)(k_top - k_lo + 1);
        src_row = ro0 + k_lo * W + (t - k_lo * a.step);
    } else {
        int64_t ro = __ldg(a.chunk_row_offsets + c0 + k_lo);
        for (int64_t k = k_lo; k <= k_hi && cover < 2; ++k) {
            const int64_t ro_next = __ldg(a.chunk_row_offsets + c0 + k + 1);
            const int64_t j = t - k * a.step;
            if (j < ro_next - ro) {  // j >= 0 by construction of k_hi
                if (cover == 0) src_row = ro + j;
                ++cover;
            }
            ro = ro_next;
        }
    }
}

// A warp owns tiles of 64 consecutive output rows, two per lane.  A lane finds the winning chunk
// of its rows (the first chunk covering each) and their coverage, loads the 20-byte source rows,
// normalises where the reference does, and parks the results in the warp's shared-memory tile;
// the tile then leaves with coalesced 16-byte stores (tile starts are 16-byte aligned).
// Every warp walks a contiguous range of tiles, so the read owning the tile start only moves
// forward: one binary search per warp, then a linear advance.
template <bool F64>
__global__ void __launch_bounds__(kWarps * 32) assemble_kernel(const AssembleArgs a)
{
    using OT = typename std::conditional<F64, double, float>::type;
    __shared__ __align__(16) OT tiles[kWarps][kTileRows * 5];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    OT *tile = tiles[warp];
    const int64_t n_tiles = (a.total_rows + kTileRows - 1) / kTileRows;
    const int64_t n_warps = (int64_t)gridDim.x * kWarps;
    const int64_t per = (n_tiles + n_warps - 1) / n_warps;
    const int64_t t_begin = ((int64_t)blockIdx.x * kWarps + warp) * per;
    int64_t t_end = t_begin + per;
    if (t_end > n_tiles) t_end = n_tiles;
    int r_lo = 0;
    if (t_begin < t_end) r_lo = find_read(a.out_row_offsets, a.n_reads, t_begin * kTileRows);
    for (int64_t tix = t_begin; tix < t_end; ++tix) {
        const int64_t g0 = tix * kTileRows;
        int64_t g_last = g0 + kTileRows - 1;
        if (g_last >= a.total_rows) g_last = a.total_rows - 1;
        while (r_lo + 1 < a.n_reads && __ldg(a.out_row_offsets + r_lo + 1) <= g0) ++r_lo;
        const bool one_read = (r_lo + 1 >= a.n_reads) || (__ldg(a.out_row_offsets + r_lo + 1) > g_last);
        const int64_t base_r = __ldg(a.out_row_offsets + r_lo);
        float x[kRowsPerLane][5];
        int cover[kRowsPerLane];
#pragma unroll
        for (int q = 0; q < kRowsPerLane; ++q) {
            const int64_t g = g0 + q * 32 + lane;
            cover[q] = 0;
            if (g <= g_last) {
                int r = r_lo;
                int64_t t = g - base_r;
                if (!one_read) {
                    r = find_read(a.out_row_offsets, a.n_reads, g);
                    t = g - __ldg(a.out_row_offsets + r);
                }
                int64_t src_row;
                locate_row(a, r, t, src_row, cover[q]);
                const float *src = a.chunks + src_row * 5;
#pragma unroll
                for (int i = 0; i < 5; ++i) x[q][i] = __ldg(src + i);
            }
        }
#pragma unroll
        for (int q = 0; q < kRowsPerLane; ++q) {
            if (g0 + q * 32 + lane <= g_last) {
                OT *o = tile + (q * 32 + lane) * 5;
                if (F64) {
                    double y[5];
#pragma unroll
                    for (int i = 0; i < 5; ++i) y[i] = (double)x[q][i];
                    if (cover[q] > 1) {
                        double nrm = 0.0;
#pragma unroll
                        for (int i = 0; i < 5; ++i) nrm = __dadd_rn(nrm, fabs(y[i]));
                        if (nrm < 10.0 * 2.220446049250313e-16) nrm = 1.0;
#pragma unroll
                        for (int i = 0; i < 5; ++i) y[i] = y[i] / nrm;
                    }
#pragma unroll
                    for (int i = 0; i < 5; ++i) o[i] = (OT)y[i];
                } else {
#pragma unroll
                    for (int i = 0; i < 5; ++i) o[i] = (OT)x[q][i];
                }
            }
        }
        __syncwarp();
        // coalesced write-back of the tile
        const int n_el = (int)(g_last - g0 + 1) * 5;
        OT *dst = (OT *)a.out + g0 * 5;
        constexpr int VEC = 16 / sizeof(OT);
        const int n_vec = n_el / VEC;
        const int4 *tv = reinterpret_cast<const int4 *>(tile);
        int4 *dv = reinterpret_cast<int4 *>(dst);
        for (int i = lane; i < n_vec; i += 32) dv[i] = tv[i];
        for (int i = n_vec * VEC + lane; i < n_el; i += 32) dst[i] = tile[i];
        __syncwarp();
    }
}

int assemble_launch(const AssembleArgs &a, bool f64, int device, cudaStream_t stream)
{
    if (a.total_rows == 0) return 0;
    DeviceInfo di;
    int rc = device_info(device, &di);
    if (rc) return rc;
    const int block = kWarps * 32;
    // one wave of resident CTAs: every warp walks one contiguous range of tiles
    int per_sm = 0;
    if (f64)
        RADIAN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, assemble_kernel<true>, block, 0));
    else
        RADIAN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, assemble_kernel<false>, block, 0));
    if (per_sm < 1) per_sm = 1;
    int64_t need = (a.total_rows + kTileRows * kWarps - 1) / (kTileRows * kWarps);
    int64_t maxg = (int64_t)di.sm_count * per_sm;
    int grid = (int)(need < maxg ? need : maxg);
    if (grid < 1) grid = 1;
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
    // sign of max_chunk_rows: positive = every chunk except each read's last has exactly that many
    // rows (the reference's windowing; enables the closed-form lookup), negative = ragged layout
    int uniform = 1;
    for (int r = 0; r < n_reads && uniform; ++r)
        for (int64_t c = read_chunk_ranges[r]; c + 1 < read_chunk_ranges[r + 1]; ++c)
            if (chunk_row_offsets[c + 1] - chunk_row_offsets[c] != maxrows) {
                uniform = 0;
                break;
            }
    if (out_max_chunk_rows) *out_max_chunk_rows = (int32_t)(uniform ? maxrows : -maxrows);
    return RADIAN_OK;
}

extern "C" int radian_assemble_batch_dev(const float *chunks, const int64_t *chunk_row_offsets,
                                         const int64_t *read_chunk_ranges, const int64_t *out_row_offsets,
                                         int n_reads, int step, int32_t max_chunk_rows,
                                         int64_t total_out_rows, void *out, int out_is_f64,
                                         radian_stream_t stream)
{
    if (n_reads < 0 || step <= 0 || total_out_rows < 0) {
        set_error("radian_assemble_batch_dev: bad arguments");
        return RADIAN_E_ARG;
    }
    if (n_reads == 0) return RADIAN_OK;
    int device = 0;
    RADIAN_CUDA(cudaGetDevice(&device));
    const int mcr = max_chunk_rows < 0 ? -max_chunk_rows : max_chunk_rows;
    AssembleArgs a{chunks, chunk_row_offsets, read_chunk_ranges, out_row_offsets, n_reads, step,
                   mcr < 1 ? 1 : mcr, out, total_out_rows, max_chunk_rows > 0};
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
    std::lock_guard<std::mutex> host_lock(host_mutex(device));
    RADIAN_CUDA(cudaSetDevice(device));
    cudaMemPool_t pool_ = nullptr;  // this library's own stream-ordered pool on the device
    {
        int krc = keep_pool(device, &pool_);
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
    TRY(cudaMallocFromPoolAsync(&d_chunks, (size_t)(in_rows ? in_rows : 1) * 20, pool_, st));
    TRY(cudaMallocFromPoolAsync(&d_cro, (size_t)(n_chunks + 1) * 8, pool_, st));
    TRY(cudaMallocFromPoolAsync(&d_rcr, (size_t)(n_reads + 1) * 8, pool_, st));
    TRY(cudaMallocFromPoolAsync(&d_oro, (size_t)(n_reads + 1) * 8, pool_, st));
    TRY(cudaMallocFromPoolAsync(&d_out, out_bytes ? out_bytes : 1, pool_, st));
    TRY(cudaMemcpyAsync(d_chunks, chunks, (size_t)in_rows * 20, cudaMemcpyHostToDevice, st));
    TRY(cudaMemcpyAsync(d_cro, chunk_row_offsets, (size_t)(n_chunks + 1) * 8, cudaMemcpyHostToDevice, st));
    TRY(cudaMemcpyAsync(d_rcr, read_chunk_ranges, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    TRY(cudaMemcpyAsync(d_oro, out_row_offsets, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    if (ret == RADIAN_OK) {
        const int mcr = maxrows < 0 ? -maxrows : maxrows;
        AssembleArgs a{d_chunks, d_cro, d_rcr, d_oro, n_reads, step, mcr < 1 ? 1 : mcr, d_out, out_rows, maxrows > 0};
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
