// Chunk-mode stitching on the GPU: what radian/sequence_assembly.py:19-48 (simple_assembly,
// add_count) followed by np.argmax + index2base (basecall.py:122-123, sequence_assembly.py:90-97)
// computes for one read, for a batch of reads.
//
// The reference walks the fragments of a read one after the other; only two things are
// sequential in that walk, and both are cheap: the running position (a prefix sum of the
// displacements) and the growth of the vote buffer.  Everything else is independent work:
//   pair_kernel    one thread per consecutive fragment pair: displacement of the first largest
//                  matching block of difflib.SequenceMatcher(None, prev, cur)
//   place_kernel   one thread per read: prefix sum of the displacements, the reference's
//                  1000-column growth rule and its IndexError, consensus length
//   vote_kernel    one thread per fragment symbol: one atomicAdd on a 4 x 16-bit packed counter
//   argmax_kernel  one thread per consensus column: first maximum, optional vote counts
//
// difflib (CPython standard library) restated: with fewer than 200 symbols in `cur` every symbol
// is in b2j, find_longest_match returns the longest common substring that ends first in `prev`
// (then first in `cur`), no block found later by the recursion is as long and starts earlier, so
// the first largest block is the top-level match.  From 200 symbols on the autojunk rule drops
// "popular" symbols from b2j and the top-level match need not be the largest block any more: those
// pairs run the complete get_matching_blocks recursion in global scratch.
#include <limits.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "internal.h"

namespace radian {

constexpr int kStitchFast = 200;  // difflib's autojunk threshold

struct Blk {
    int i, j, k;
};

// find_longest_match(alo, ahi, blo, bhi); rows: two int arrays of bhi + 2 entries, indexed j + 1
template <typename RowT>
__device__ Blk longest_match(const uint8_t *a, const uint8_t *b, int alo, int ahi, int blo, int bhi,
                             unsigned popular_mask, RowT *row0, RowT *row1)
{
    int besti = alo, bestj = blo, bestsize = 0;
    RowT *old = row0, *cur = row1;
    for (int j = blo; j <= bhi; ++j) old[j] = 0;
    for (int i = alo; i < ahi; ++i) {
        const int ai = a[i];
        cur[blo] = 0;
        const bool usable = !((popular_mask >> ai) & 1u);
        for (int j = blo; j < bhi; ++j) {
            int k = 0;
            if (usable && b[j] == ai) {
                k = (int)old[j] + 1;  // j2len.get(j - 1, 0) + 1
                if (k > bestsize) {
                    besti = i - k + 1;
                    bestj = j - k + 1;
                    bestsize = k;
                }
            }
            cur[j + 1] = (RowT)k;
        }
        RowT *t = old;
        old = cur;
        cur = t;
    }
    // "extend the best by non-junk elements on each end" (popular symbols included)
    while (besti > alo && bestj > blo && a[besti - 1] == b[bestj - 1]) {
        --besti;
        --bestj;
        ++bestsize;
    }
    while (besti + bestsize < ahi && bestj + bestsize < bhi && a[besti + bestsize] == b[bestj + bestsize]) ++bestsize;
    return Blk{besti, bestj, bestsize};
}

// the complete get_matching_blocks + max(key=size); scratch: ints, see pair_scratch_ints()
__device__ int disp_general(const uint8_t *a, int la, const uint8_t *b, int lb, int *scratch)
{
    unsigned popular = 0;
    {
        int cnt[4] = {0, 0, 0, 0};
        for (int j = 0; j < lb; ++j) cnt[b[j]]++;
        for (int c = 0; c < 4; ++c)
            if (cnt[c] > lb / 100 + 1) popular |= 1u << c;
    }
    const int nmax = (la < lb ? la : lb) + 2;
    Blk *blocks = reinterpret_cast<Blk *>(scratch);
    int *queue = scratch + 3 * nmax;
    int *row0 = queue + 4 * (2 * nmax + 2);
    int *row1 = row0 + lb + 2;
    int nb = 0, qh = 0, qt = 1;
    queue[0] = 0, queue[1] = la, queue[2] = 0, queue[3] = lb;
    while (qh < qt) {
        const int alo = queue[4 * qh], ahi = queue[4 * qh + 1], blo = queue[4 * qh + 2], bhi = queue[4 * qh + 3];
        ++qh;
        const Blk x = longest_match<int>(a, b, alo, ahi, blo, bhi, popular, row0, row1);
        if (x.k) {
            blocks[nb++] = x;
            if (alo < x.i && blo < x.j) {
                queue[4 * qt] = alo, queue[4 * qt + 1] = x.i, queue[4 * qt + 2] = blo, queue[4 * qt + 3] = x.j;
                ++qt;
            }
            if (x.i + x.k < ahi && x.j + x.k < bhi) {
                queue[4 * qt] = x.i + x.k, queue[4 * qt + 1] = ahi, queue[4 * qt + 2] = x.j + x.k, queue[4 * qt + 3] = bhi;
                ++qt;
            }
        }
    }
    // sort by (i, j): blocks never share a start in `a`, insertion sort on i
    for (int n = 1; n < nb; ++n) {
        const Blk x = blocks[n];
        int m = n - 1;
        while (m >= 0 && (blocks[m].i > x.i || (blocks[m].i == x.i && blocks[m].j > x.j))) {
            blocks[m + 1] = blocks[m];
            --m;
        }
        blocks[m + 1] = x;
    }
    // collapse adjacent blocks; first block of maximal size; else the (la, lb, 0) sentinel
    int bi = la, bj = lb, bk = 0;
    int i1 = 0, j1 = 0, k1 = 0;
    for (int n = 0; n <= nb; ++n) {
        const bool last = (n == nb);
        if (!last && i1 + k1 == blocks[n].i && j1 + k1 == blocks[n].j) {
            k1 += blocks[n].k;
        } else {
            if (k1 > bk) bi = i1, bj = j1, bk = k1;
            if (!last) i1 = blocks[n].i, j1 = blocks[n].j, k1 = blocks[n].k;
        }
    }
    return bi - bj;
}

__host__ __device__ inline int64_t pair_scratch_ints(int64_t la, int64_t lb)
{
    const int64_t nmax = (la < lb ? la : lb) + 2;
    return 3 * nmax + 4 * (2 * nmax + 2) + 2 * (lb + 2);
}

// which read owns fragment f (one thread per read fills its range)
__global__ void owner_kernel(const int64_t *__restrict__ read_frag, int n_reads, int32_t *__restrict__ frag_read)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    for (int64_t f = read_frag[r]; f < read_frag[r + 1]; ++f) frag_read[f] = r;
}

// disp[f] = block[0] - block[1] for fragment f against its predecessor (0 for a read's first).
// Fragments are given by start and length, so the decoder's output slots can be used as they are.
// Pairs with 200 symbols or more take their scratch from a pool (bump allocation); a pair that
// finds it empty marks the read (disp = INT_MIN).
__global__ void pair_kernel(const uint8_t *__restrict__ sym, const int64_t *__restrict__ frag_start,
                            const int64_t *__restrict__ frag_len, const int64_t *__restrict__ read_frag,
                            const int32_t *__restrict__ frag_read, int64_t n_frags, int *__restrict__ pool,
                            unsigned long long *__restrict__ pool_used, int64_t pool_ints,
                            int32_t *__restrict__ disp)
{
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_frags) return;
    if (f == read_frag[frag_read[f]]) {
        disp[f] = 0;
        return;
    }
    const uint8_t *a = sym + frag_start[f - 1];
    const uint8_t *b = sym + frag_start[f];
    const int la = (int)frag_len[f - 1];
    const int lb = (int)frag_len[f];
    if (lb < kStitchFast) {
        uint16_t row0[kStitchFast + 2], row1[kStitchFast + 2];  // a match is at most lb < 200 long
        const Blk x = longest_match<uint16_t>(a, b, 0, la, 0, lb, 0u, row0, row1);
        disp[f] = x.k ? x.i - x.j : la - lb;  // no common symbol: only the (la, lb, 0) sentinel
    } else {
        const int64_t need = pair_scratch_ints(la, lb);
        const unsigned long long at = atomicAdd(pool_used, (unsigned long long)need);
        disp[f] = (int64_t)at + need <= pool_ints ? disp_general(a, la, b, lb, pool + at) : INT_MIN;
    }
}

// per read: running position, growth rule, IndexError, consensus length
__global__ void place_kernel(const int64_t *__restrict__ frag_len, const int64_t *__restrict__ read_frag,
                             const int64_t *__restrict__ out_offsets, int n_reads,
                             const int32_t *__restrict__ disp, int64_t *__restrict__ start,
                             int64_t *__restrict__ out_len, int32_t *__restrict__ status)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    int64_t census = 1000, pos = 0, length = 0;
    int st = RADIAN_READ_OK;
    const int64_t f0 = read_frag[r], f1 = read_frag[r + 1];
    if (f1 - f0 > 65535) st = RADIAN_READ_SEQ_OVERFLOW;  // the packed vote counters are 16 bits wide
    for (int64_t f = f0; f < f1; ++f) {
        const int64_t len = frag_len[f];
        int64_t s = 0;
        if (f > f0) {
            const int64_t d = disp[f];
            if (d == INT_MIN && st == RADIAN_READ_OK) st = RADIAN_READ_TRIE_OVERFLOW;  // scratch pool exhausted
            if (d + pos + len > census) census += 1000;  // sequence_assembly.py:33-36, one step only
            s = pos + d;
            pos += d;
            if (pos + len > length) length = pos + len;
        }
        start[f] = s;  // may be negative: add_count drops the first -s symbols (sequence_assembly.py:43-45)
        const int64_t skip = s < 0 ? -s : 0;
        const int64_t s0 = s < 0 ? 0 : s;
        if (len > skip && s0 + (len - skip) > census && st == RADIAN_READ_OK) st = RADIAN_READ_INDEX_ERROR;
    }
    if (length > census) length = census;
    if (st == RADIAN_READ_OK && length > out_offsets[r + 1] - out_offsets[r]) st = RADIAN_READ_SEQ_OVERFLOW;
    out_len[r] = st == RADIAN_READ_OK ? length : 0;
    status[r] = st;
}

// votes: one 64-bit word per consensus column, 16 bits per base
__global__ void vote_kernel(const uint8_t *__restrict__ sym, const int64_t *__restrict__ frag_start,
                            const int64_t *__restrict__ frag_len, const int32_t *__restrict__ frag_read,
                            int64_t n_frags, const int64_t *__restrict__ start, const int64_t *__restrict__ out_len,
                            const int64_t *__restrict__ out_offsets, unsigned long long *__restrict__ votes)
{
    const int64_t f = blockIdx.x;
    if (f >= n_frags) return;
    const int r = frag_read[f];
    const int64_t length = out_len[r];
    const int64_t s = start[f];
    const int64_t len = frag_len[f];
    const uint8_t *p = sym + frag_start[f];
    unsigned long long *v = votes + out_offsets[r];
    for (int64_t i = threadIdx.x; i < len; i += blockDim.x) {
        const int64_t col = s + i;  // s < 0: the first -s symbols fall off
        if (col >= 0 && col < length) atomicAdd(&v[col], 1ull << (16 * p[i]));
    }
}

// one block per read: first maximum of every consensus column, optional vote counts
__global__ void argmax_kernel(const unsigned long long *__restrict__ votes, const int64_t *__restrict__ out_offsets,
                              const int64_t *__restrict__ out_len, int n_reads, uint8_t *__restrict__ seq,
                              int32_t *__restrict__ counts)
{
    for (int r = blockIdx.x; r < n_reads; r += gridDim.x) {
        const int64_t o = out_offsets[r], n = out_len[r];
        for (int64_t c = threadIdx.x; c < n; c += blockDim.x) {
            const unsigned long long w = votes[o + c];
            int best = 0, bv = (int)(w & 0xffff);
#pragma unroll
            for (int s = 1; s < 4; ++s) {
                const int x = (int)((w >> (16 * s)) & 0xffff);
                if (x > bv) {  // np.argmax: first maximum
                    bv = x;
                    best = s;
                }
            }
            seq[o + c] = (uint8_t)best;
            if (counts) {
#pragma unroll
                for (int s = 0; s < 4; ++s) counts[(o + c) * 4 + s] = (int)((w >> (16 * s)) & 0xffff);
            }
        }
    }
}

struct StitchWs {
    int32_t *frag_read, *disp;
    int64_t *start;
    unsigned long long *votes, *pool_used;
    int *pool;
    int64_t pool_ints;
};

static size_t stitch_ws_layout(int64_t n_frags, int64_t total_slots, int64_t pool_ints, void *base, StitchWs *ws)
{
    size_t o = 0;
    auto take = [&](size_t bytes) {
        const size_t at = o;
        o += (bytes + 255) & ~(size_t)255;
        return at;
    };
    const size_t o_used = take(8), o_votes = take((size_t)total_slots * 8), o_start = take((size_t)n_frags * 8),
                 o_read = take((size_t)n_frags * 4), o_disp = take((size_t)n_frags * 4),
                 o_pool = take((size_t)pool_ints * 4);
    if (ws) {
        char *b = (char *)base;
        ws->pool_used = (unsigned long long *)(b + o_used);
        ws->votes = (unsigned long long *)(b + o_votes);
        ws->start = (int64_t *)(b + o_start);
        ws->frag_read = (int32_t *)(b + o_read);
        ws->disp = (int32_t *)(b + o_disp);
        ws->pool = (int *)(b + o_pool);
        ws->pool_ints = pool_ints;
    }
    return o;
}

}  // namespace radian

using namespace radian;

extern "C" size_t radian_stitch_workspace_bytes(int64_t n_frags, int64_t total_slots, int64_t long_pair_scratch_ints)
{
    if (n_frags < 0 || total_slots < 0 || long_pair_scratch_ints < 0) return 0;
    return stitch_ws_layout(n_frags, total_slots, long_pair_scratch_ints, nullptr, nullptr);
}

extern "C" int radian_stitch_batch_dev(const uint8_t *frag_sym, const int64_t *frag_start, const int64_t *frag_len,
                                       const int64_t *read_frag_ranges, int n_reads, int64_t n_frags,
                                       uint8_t *out_seq, const int64_t *out_offsets, int64_t total_slots,
                                       int64_t *out_len, int32_t *out_status, int32_t *out_votes,
                                       int64_t long_pair_scratch_ints, void *workspace, size_t workspace_bytes,
                                       radian_stream_t stream)
{
    if (n_reads < 0 || n_frags < 0 || !read_frag_ranges || !out_offsets || !out_len || !out_status ||
        (n_frags > 0 && (!frag_start || !frag_len))) {
        set_error("radian_stitch_batch_dev: null argument");
        return RADIAN_E_ARG;
    }
    if (n_reads == 0) return RADIAN_OK;
    StitchWs ws;
    const size_t need = stitch_ws_layout(n_frags, total_slots, long_pair_scratch_ints, workspace, &ws);
    if (!workspace || workspace_bytes < need) {
        set_error("radian_stitch_batch_dev: workspace of %zu bytes needed, %zu given", need, workspace_bytes);
        return RADIAN_E_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    RADIAN_CUDA(cudaMemsetAsync(ws.pool_used, 0, 8, st));
    if (total_slots > 0) RADIAN_CUDA(cudaMemsetAsync(ws.votes, 0, (size_t)total_slots * 8, st));
    const unsigned rb = (unsigned)((n_reads + 127) / 128);
    owner_kernel<<<rb, 128, 0, st>>>(read_frag_ranges, n_reads, ws.frag_read);
    if (n_frags > 0)
        pair_kernel<<<(unsigned)((n_frags + 127) / 128), 128, 0, st>>>(frag_sym, frag_start, frag_len, read_frag_ranges,
                                                                         ws.frag_read, n_frags, ws.pool, ws.pool_used,
                                                                         ws.pool_ints, ws.disp);
    place_kernel<<<rb, 128, 0, st>>>(frag_len, read_frag_ranges, out_offsets, n_reads, ws.disp, ws.start, out_len,
                                     out_status);
    if (n_frags > 0 && total_slots > 0) {
        vote_kernel<<<(unsigned)n_frags, 64, 0, st>>>(frag_sym, frag_start, frag_len, ws.frag_read, n_frags, ws.start,
                                                      out_len, out_offsets, ws.votes);
        int device = 0;
        RADIAN_CUDA(cudaGetDevice(&device));
        DeviceInfo di;
        int rc = device_info(device, &di);
        if (rc) return rc;
        const int grid = n_reads < di.sm_count * 16 ? n_reads : di.sm_count * 16;
        argmax_kernel<<<grid, 128, 0, st>>>(ws.votes, out_offsets, out_len, n_reads, out_seq, out_votes);
    }
    RADIAN_CUDA(cudaGetLastError());
    return RADIAN_OK;
}

extern "C" int radian_stitch_batch_host(const uint8_t *frag_sym, const int64_t *frag_offsets,
                                        const int64_t *read_frag_ranges, int n_reads, uint8_t *out_seq,
                                        const int64_t *out_offsets, int64_t *out_len, int32_t *out_status,
                                        int32_t *out_votes, int device)
{
    if (n_reads < 0 || !frag_offsets || !read_frag_ranges || !out_offsets || !out_len || !out_status) {
        set_error("radian_stitch_batch_host: null argument");
        return RADIAN_E_ARG;
    }
    if (n_reads == 0) return RADIAN_OK;
    if (radian_device_count() <= device || device < 0) {
        set_error("radian_stitch_batch_host: CUDA device %d not available (no CPU fallback exists)", device);
        return RADIAN_E_CUDA;
    }
    const int64_t n_frags = read_frag_ranges[n_reads];
    if (read_frag_ranges[0] != 0 || n_frags < 0 || out_offsets[0] != 0) {
        set_error("radian_stitch_batch_host: read_frag_ranges and out_offsets must start at 0");
        return RADIAN_E_ARG;
    }
    for (int r = 0; r < n_reads; ++r)
        if (read_frag_ranges[r + 1] < read_frag_ranges[r] || out_offsets[r + 1] < out_offsets[r]) {
            set_error("radian_stitch_batch_host: offsets not monotone at read %d", r);
            return RADIAN_E_ARG;
        }
    const int64_t n_sym = n_frags ? frag_offsets[n_frags] : 0;
    int64_t pool_ints = 0;
    std::vector<int64_t> flen((size_t)(n_frags ? n_frags : 1));
    for (int64_t f = 0; f < n_frags; ++f) {
        if (frag_offsets[f + 1] < frag_offsets[f]) {
            set_error("radian_stitch_batch_host: fragment offsets not monotone at fragment %lld", (long long)f);
            return RADIAN_E_ARG;
        }
        flen[f] = frag_offsets[f + 1] - frag_offsets[f];
        if (f > 0 && flen[f] >= kStitchFast) pool_ints += pair_scratch_ints(flen[f - 1], flen[f]);
    }
    for (int64_t i = 0; i < n_sym; ++i)
        if (frag_sym[i] > 3) {
            set_error("radian_stitch_batch_host: symbol %d at %lld is not a base (KeyError in add_count, "
                      "sequence_assembly.py:42)", (int)frag_sym[i], (long long)i);
            return RADIAN_E_CONTEXT;
        }
    if (n_frags == 0) {
        for (int r = 0; r < n_reads; ++r) out_len[r] = 0, out_status[r] = RADIAN_READ_OK;
        return RADIAN_OK;
    }
    std::lock_guard<std::mutex> host_lock(host_mutex(device));
    RADIAN_CUDA(cudaSetDevice(device));
    cudaMemPool_t pool_ = nullptr;  // this library's own stream-ordered pool on the device
    {
        int krc = keep_pool(device, &pool_);
        if (krc) return krc;
    }
    static const bool trace = getenv("RADIAN_TRACE") != nullptr;
    const int64_t total_slots = out_offsets[n_reads];
    const size_t ws_bytes = radian_stitch_workspace_bytes(n_frags, total_slots, pool_ints);
    cudaStream_t st = nullptr;
    RADIAN_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    cudaEvent_t evt[2] = {nullptr, nullptr};
    if (trace)
        for (auto &x : evt) RADIAN_CUDA(cudaEventCreate(&x));
    uint8_t *d_sym = nullptr, *d_seq = nullptr;
    int64_t *d_fstart = nullptr, *d_flen = nullptr, *d_rfr = nullptr, *d_ooff = nullptr, *d_len = nullptr;
    int32_t *d_status = nullptr, *d_counts = nullptr;
    void *d_ws = nullptr;
    int ret = RADIAN_OK;
    cudaError_t e;
#define TRY(x)                                   \
    if (ret == RADIAN_OK && (e = (x)) != cudaSuccess) ret = cuda_fail(e, #x)
    TRY(cudaMallocFromPoolAsync(&d_sym, (size_t)(n_sym ? n_sym : 1), pool_, st));
    TRY(cudaMallocFromPoolAsync(&d_seq, (size_t)(total_slots ? total_slots : 1), pool_, st));
    TRY(cudaMallocFromPoolAsync(&d_fstart, (size_t)n_frags * 8, pool_, st));
    TRY(cudaMallocFromPoolAsync(&d_flen, (size_t)n_frags * 8, pool_, st));
    TRY(cudaMallocFromPoolAsync(&d_rfr, (size_t)(n_reads + 1) * 8, pool_, st));
    TRY(cudaMallocFromPoolAsync(&d_ooff, (size_t)(n_reads + 1) * 8, pool_, st));
    TRY(cudaMallocFromPoolAsync(&d_len, (size_t)n_reads * 8, pool_, st));
    TRY(cudaMallocFromPoolAsync(&d_status, (size_t)n_reads * 4, pool_, st));
    TRY(cudaMallocFromPoolAsync(&d_ws, ws_bytes, pool_, st));
    if (out_votes) TRY(cudaMallocFromPoolAsync(&d_counts, (size_t)(total_slots ? total_slots : 1) * 16, pool_, st));
    if (n_sym) TRY(cudaMemcpyAsync(d_sym, frag_sym, (size_t)n_sym, cudaMemcpyHostToDevice, st));
    TRY(cudaMemcpyAsync(d_fstart, frag_offsets, (size_t)n_frags * 8, cudaMemcpyHostToDevice, st));
    TRY(cudaMemcpyAsync(d_flen, flen.data(), (size_t)n_frags * 8, cudaMemcpyHostToDevice, st));
    TRY(cudaMemcpyAsync(d_rfr, read_frag_ranges, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    TRY(cudaMemcpyAsync(d_ooff, out_offsets, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    if (trace && ret == RADIAN_OK) cudaEventRecord(evt[0], st);
    if (ret == RADIAN_OK)
        ret = radian_stitch_batch_dev(d_sym, d_fstart, d_flen, d_rfr, n_reads, n_frags, d_seq, d_ooff, total_slots,
                                      d_len, d_status, d_counts, pool_ints, d_ws, ws_bytes, st);
    if (trace && ret == RADIAN_OK) cudaEventRecord(evt[1], st);
    TRY(cudaMemcpyAsync(out_len, d_len, (size_t)n_reads * 8, cudaMemcpyDeviceToHost, st));
    TRY(cudaMemcpyAsync(out_status, d_status, (size_t)n_reads * 4, cudaMemcpyDeviceToHost, st));
    if (total_slots) TRY(cudaMemcpyAsync(out_seq, d_seq, (size_t)total_slots, cudaMemcpyDeviceToHost, st));
    if (out_votes && total_slots)
        TRY(cudaMemcpyAsync(out_votes, d_counts, (size_t)total_slots * 16, cudaMemcpyDeviceToHost, st));
    TRY(cudaStreamSynchronize(st));
#undef TRY
    if (trace && ret == RADIAN_OK) {
        float ms = 0;
        cudaEventElapsedTime(&ms, evt[0], evt[1]);
        fprintf(stderr, "[radian] stitch: %d reads, %lld fragments, %lld symbols | kernels %.3f ms\n", n_reads,
                (long long)n_frags, (long long)n_sym, ms);
    }
    if (trace)
        for (auto &x : evt) cudaEventDestroy(x);
    void *frees[] = {d_sym, d_seq, d_fstart, d_flen, d_rfr, d_ooff, d_len, d_status, d_ws, d_counts};
    for (void *p : frees)
        if (p) cudaFreeAsync(p, st);
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    if (ret != RADIAN_OK) return ret;
    for (int r = 0; r < n_reads; ++r)
        if (out_status[r] != RADIAN_READ_OK) {
            if (out_status[r] == RADIAN_READ_SEQ_OVERFLOW) {
                set_error("radian_stitch_batch_host: read %d: output slot smaller than the consensus, or more than "
                          "65535 fragments (the sum of a read's fragment lengths always suffices as a slot)", r);
                return RADIAN_E_ARG;
            }
            set_error("radian_stitch_batch_host: read %d failed with status %d%s", r, out_status[r],
                      out_status[r] == RADIAN_READ_INDEX_ERROR
                          ? " (a fragment does not fit the reference's vote buffer: IndexError in add_count, "
                            "sequence_assembly.py:47)"
                          : "");
            return out_status[r] == RADIAN_READ_INDEX_ERROR ? RADIAN_E_GAP : RADIAN_E_READ;
        }
    return RADIAN_OK;
}
