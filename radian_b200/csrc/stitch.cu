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

// disp[f] = block[0] - block[1] for fragment f against its predecessor (0 for a read's first)
__global__ void pair_kernel(const uint8_t *__restrict__ sym, const int64_t *__restrict__ frag_off,
                            const uint8_t *__restrict__ is_first, int64_t n_frags,
                            const int64_t *__restrict__ scratch_off, int *__restrict__ scratch,
                            int32_t *__restrict__ disp)
{
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_frags) return;
    if (is_first[f]) {
        disp[f] = 0;
        return;
    }
    const uint8_t *a = sym + frag_off[f - 1];
    const uint8_t *b = sym + frag_off[f];
    const int la = (int)(frag_off[f] - frag_off[f - 1]);
    const int lb = (int)(frag_off[f + 1] - frag_off[f]);
    if (lb < kStitchFast) {
        uint16_t row0[kStitchFast + 2], row1[kStitchFast + 2];  // a match is at most lb < 200 long
        const Blk x = longest_match<uint16_t>(a, b, 0, la, 0, lb, 0u, row0, row1);
        disp[f] = x.k ? x.i - x.j : la - lb;  // no common symbol: only the (la, lb, 0) sentinel
    } else {
        disp[f] = disp_general(a, la, b, lb, scratch + scratch_off[f]);
    }
}

// per read: running position, growth rule, IndexError, consensus length
__global__ void place_kernel(const int64_t *__restrict__ frag_off, const int64_t *__restrict__ read_frag,
                             int n_reads, const int32_t *__restrict__ disp, int64_t *__restrict__ start,
                             int64_t *__restrict__ out_len, int32_t *__restrict__ status)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    int64_t census = 1000, pos = 0, length = 0;
    int st = RADIAN_READ_OK;
    const int64_t f0 = read_frag[r], f1 = read_frag[r + 1];
    if (f1 - f0 > 65535) st = RADIAN_READ_SEQ_OVERFLOW;  // the packed vote counters are 16 bits wide
    for (int64_t f = f0; f < f1; ++f) {
        const int64_t len = frag_off[f + 1] - frag_off[f];
        int64_t s = 0;
        if (f > f0) {
            const int64_t d = disp[f];
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
    out_len[r] = st == RADIAN_READ_OK ? length : 0;
    status[r] = st;
}

// votes: one 64-bit word per consensus column, 16 bits per base
__global__ void vote_kernel(const uint8_t *__restrict__ sym, const int64_t *__restrict__ frag_off,
                            const int32_t *__restrict__ frag_read, int64_t n_frags,
                            const int64_t *__restrict__ start, const int64_t *__restrict__ out_len,
                            const int64_t *__restrict__ col_off, unsigned long long *__restrict__ votes)
{
    const int64_t f = blockIdx.x;
    if (f >= n_frags) return;
    const int r = frag_read[f];
    const int64_t length = out_len[r];
    const int64_t s = start[f];
    const int64_t len = frag_off[f + 1] - frag_off[f];
    const uint8_t *p = sym + frag_off[f];
    unsigned long long *v = votes + col_off[r];
    for (int64_t i = threadIdx.x; i < len; i += blockDim.x) {
        const int64_t col = s + i;  // s < 0: the first -s symbols fall off
        if (col >= 0 && col < length) atomicAdd(&v[col], 1ull << (16 * p[i]));
    }
}

__global__ void argmax_kernel(const unsigned long long *__restrict__ votes, int64_t n_cols,
                              uint8_t *__restrict__ seq, int32_t *__restrict__ counts)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cols) return;
    const unsigned long long w = votes[c];
    int best = 0, bv = (int)(w & 0xffff);
#pragma unroll
    for (int s = 1; s < 4; ++s) {
        const int x = (int)((w >> (16 * s)) & 0xffff);
        if (x > bv) {  // np.argmax: first maximum
            bv = x;
            best = s;
        }
    }
    seq[c] = (uint8_t)best;
    if (counts) {
#pragma unroll
        for (int s = 0; s < 4; ++s) counts[c * 4 + s] = (int)((w >> (16 * s)) & 0xffff);
    }
}

}  // namespace radian

using namespace radian;

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
    if (read_frag_ranges[0] != 0 || n_frags < 0) {
        set_error("radian_stitch_batch_host: read_frag_ranges must start at 0");
        return RADIAN_E_ARG;
    }
    for (int r = 0; r < n_reads; ++r)
        if (read_frag_ranges[r + 1] < read_frag_ranges[r] || out_offsets[r + 1] < out_offsets[r]) {
            set_error("radian_stitch_batch_host: offsets not monotone at read %d", r);
            return RADIAN_E_ARG;
        }
    const int64_t n_sym = n_frags ? frag_offsets[n_frags] : 0;
    for (int64_t f = 0; f < n_frags; ++f) {
        if (frag_offsets[f + 1] < frag_offsets[f]) {
            set_error("radian_stitch_batch_host: fragment offsets not monotone at fragment %lld", (long long)f);
            return RADIAN_E_ARG;
        }
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
    RADIAN_CUDA(cudaSetDevice(device));
    {
        int krc = keep_pool(device);
        if (krc) return krc;
    }
    // host-side plan: which fragment starts a read, which read owns it, scratch of the long pairs
    std::vector<uint8_t> is_first((size_t)n_frags, 0);
    std::vector<int32_t> frag_read((size_t)n_frags, 0);
    std::vector<int64_t> scratch_off((size_t)n_frags, 0);
    int64_t scratch_ints = 0;
    for (int r = 0; r < n_reads; ++r)
        for (int64_t f = read_frag_ranges[r]; f < read_frag_ranges[r + 1]; ++f) {
            is_first[f] = (f == read_frag_ranges[r]);
            frag_read[f] = r;
            const int64_t lb = frag_offsets[f + 1] - frag_offsets[f];
            if (!is_first[f] && lb >= kStitchFast) {
                scratch_off[f] = scratch_ints;
                scratch_ints += pair_scratch_ints(frag_offsets[f] - frag_offsets[f - 1], lb);
            }
        }
    static const bool trace = getenv("RADIAN_TRACE") != nullptr;
    cudaStream_t st = nullptr;
    RADIAN_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    cudaEvent_t evt[4] = {nullptr, nullptr, nullptr, nullptr};
    if (trace)
        for (auto &x : evt) RADIAN_CUDA(cudaEventCreate(&x));
    uint8_t *d_sym = nullptr, *d_first = nullptr, *d_seq = nullptr;
    int64_t *d_foff = nullptr, *d_rfr = nullptr, *d_soff = nullptr, *d_start = nullptr, *d_len = nullptr, *d_col = nullptr;
    int32_t *d_fread = nullptr, *d_disp = nullptr, *d_status = nullptr, *d_counts = nullptr;
    int *d_scratch = nullptr;
    unsigned long long *d_votes = nullptr;
    int ret = RADIAN_OK;
    cudaError_t e;
#define TRY(x)                                   \
    if (ret == RADIAN_OK && (e = (x)) != cudaSuccess) ret = cuda_fail(e, #x)
    TRY(cudaMallocAsync(&d_sym, (size_t)(n_sym ? n_sym : 1), st));
    TRY(cudaMallocAsync(&d_first, (size_t)n_frags, st));
    TRY(cudaMallocAsync(&d_foff, (size_t)(n_frags + 1) * 8, st));
    TRY(cudaMallocAsync(&d_rfr, (size_t)(n_reads + 1) * 8, st));
    TRY(cudaMallocAsync(&d_soff, (size_t)n_frags * 8, st));
    TRY(cudaMallocAsync(&d_start, (size_t)n_frags * 8, st));
    TRY(cudaMallocAsync(&d_fread, (size_t)n_frags * 4, st));
    TRY(cudaMallocAsync(&d_disp, (size_t)n_frags * 4, st));
    TRY(cudaMallocAsync(&d_len, (size_t)n_reads * 8, st));
    TRY(cudaMallocAsync(&d_col, (size_t)(n_reads + 1) * 8, st));
    TRY(cudaMallocAsync(&d_status, (size_t)n_reads * 4, st));
    TRY(cudaMallocAsync(&d_scratch, (size_t)(scratch_ints ? scratch_ints : 1) * 4, st));
    if (n_sym) TRY(cudaMemcpyAsync(d_sym, frag_sym, (size_t)n_sym, cudaMemcpyHostToDevice, st));
    TRY(cudaMemcpyAsync(d_first, is_first.data(), (size_t)n_frags, cudaMemcpyHostToDevice, st));
    TRY(cudaMemcpyAsync(d_foff, frag_offsets, (size_t)(n_frags + 1) * 8, cudaMemcpyHostToDevice, st));
    TRY(cudaMemcpyAsync(d_rfr, read_frag_ranges, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
    TRY(cudaMemcpyAsync(d_soff, scratch_off.data(), (size_t)n_frags * 8, cudaMemcpyHostToDevice, st));
    TRY(cudaMemcpyAsync(d_fread, frag_read.data(), (size_t)n_frags * 4, cudaMemcpyHostToDevice, st));
    if (ret == RADIAN_OK) {
        if (trace) cudaEventRecord(evt[0], st);
        pair_kernel<<<(unsigned)((n_frags + 127) / 128), 128, 0, st>>>(d_sym, d_foff, d_first, n_frags, d_soff,
                                                                         d_scratch, d_disp);
        place_kernel<<<(unsigned)((n_reads + 127) / 128), 128, 0, st>>>(d_foff, d_rfr, n_reads, d_disp, d_start,
                                                                          d_len, d_status);
        if (trace) cudaEventRecord(evt[1], st);
        TRY(cudaGetLastError());
    }
    TRY(cudaMemcpyAsync(out_len, d_len, (size_t)n_reads * 8, cudaMemcpyDeviceToHost, st));
    TRY(cudaMemcpyAsync(out_status, d_status, (size_t)n_reads * 4, cudaMemcpyDeviceToHost, st));
    TRY(cudaStreamSynchronize(st));
    // consensus columns of all reads back to back
    std::vector<int64_t> col((size_t)n_reads + 1, 0);
    bool slot_small = false;
    if (ret == RADIAN_OK) {
        for (int r = 0; r < n_reads; ++r) {
            col[r + 1] = col[r] + out_len[r];
            if (out_len[r] > out_offsets[r + 1] - out_offsets[r]) slot_small = true;
        }
        if (slot_small) {
            set_error("radian_stitch_batch_host: an output slot is smaller than the consensus "
                      "(the sum of a read's fragment lengths always suffices)");
            ret = RADIAN_E_ARG;
        }
    }
    const int64_t n_cols = col[n_reads];
    if (ret == RADIAN_OK && n_cols > 0) {
        TRY(cudaMallocAsync(&d_votes, (size_t)n_cols * 8, st));
        TRY(cudaMallocAsync(&d_seq, (size_t)n_cols, st));
        if (out_votes) TRY(cudaMallocAsync(&d_counts, (size_t)n_cols * 16, st));
        TRY(cudaMemsetAsync(d_votes, 0, (size_t)n_cols * 8, st));
        TRY(cudaMemcpyAsync(d_col, col.data(), (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, st));
        if (ret == RADIAN_OK) {
            if (trace) cudaEventRecord(evt[2], st);
            vote_kernel<<<(unsigned)n_frags, 64, 0, st>>>(d_sym, d_foff, d_fread, n_frags, d_start, d_len, d_col,
                                                          d_votes);
            argmax_kernel<<<(unsigned)((n_cols + 255) / 256), 256, 0, st>>>(d_votes, n_cols, d_seq, d_counts);
            if (trace) cudaEventRecord(evt[3], st);
            TRY(cudaGetLastError());
        }
        std::vector<uint8_t> h_seq((size_t)n_cols);
        std::vector<int32_t> h_counts(out_votes ? (size_t)n_cols * 4 : 0);
        TRY(cudaMemcpyAsync(h_seq.data(), d_seq, (size_t)n_cols, cudaMemcpyDeviceToHost, st));
        if (out_votes) TRY(cudaMemcpyAsync(h_counts.data(), d_counts, (size_t)n_cols * 16, cudaMemcpyDeviceToHost, st));
        TRY(cudaStreamSynchronize(st));
        if (ret == RADIAN_OK)
            for (int r = 0; r < n_reads; ++r) {
                if (out_len[r] > 0) memcpy(out_seq + out_offsets[r], h_seq.data() + col[r], (size_t)out_len[r]);
                if (out_votes && out_len[r] > 0)
                    memcpy(out_votes + out_offsets[r] * 4, h_counts.data() + col[r] * 4, (size_t)out_len[r] * 16);
            }
    }
#undef TRY
    if (trace && ret == RADIAN_OK && n_cols > 0) {
        float ms_pair = 0, ms_vote = 0;
        cudaEventElapsedTime(&ms_pair, evt[0], evt[1]);
        cudaEventElapsedTime(&ms_vote, evt[2], evt[3]);
        fprintf(stderr, "[radian] stitch: %d reads, %lld fragments, %lld symbols, %lld columns | pair+place %.3f ms, "
                "vote+argmax %.3f ms\n", n_reads, (long long)n_frags, (long long)n_sym, (long long)n_cols, ms_pair, ms_vote);
    }
    if (trace)
        for (auto &x : evt) cudaEventDestroy(x);
    void *frees[] = {d_sym, d_first, d_foff, d_rfr, d_soff, d_start, d_fread, d_disp, d_len, d_col, d_status,
                     d_scratch, d_votes, d_seq, d_counts};
    for (void *p : frees)
        if (p) cudaFreeAsync(p, st);
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    if (ret != RADIAN_OK) return ret;
    for (int r = 0; r < n_reads; ++r)
        if (out_status[r] != RADIAN_READ_OK) {
            set_error("radian_stitch_batch_host: read %d failed with status %d%s", r, out_status[r],
                      out_status[r] == RADIAN_READ_INDEX_ERROR
                          ? " (a fragment does not fit the reference's vote buffer: IndexError in add_count, "
                            "sequence_assembly.py:47)"
                          : "");
            return out_status[r] == RADIAN_READ_INDEX_ERROR ? RADIAN_E_GAP : RADIAN_E_READ;
        }
    return RADIAN_OK;
}
