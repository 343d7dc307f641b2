// C ABI of libradian_b200.so: error reporting, the resident RNA table, decode entry points.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <thread>
#include <vector>

#include "internal.h"

namespace radian {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what)
{
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return RADIAN_E_CUDA;
}

int device_info(int device, DeviceInfo *out)
{
    static std::mutex mu;
    static DeviceInfo cache[64];
    static bool have[64] = {false};
    if (device < 0 || device >= 64) {
        set_error("bad device index %d", device);
        return RADIAN_E_ARG;
    }
    std::lock_guard<std::mutex> lk(mu);
    if (!have[device]) {
        RADIAN_CUDA(cudaDeviceGetAttribute(&cache[device].sm_count, cudaDevAttrMultiProcessorCount, device));
        RADIAN_CUDA(cudaDeviceGetAttribute(&cache[device].max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        have[device] = true;
    }
    *out = cache[device];
    return 0;
}

// The _host entry points allocate with cudaMallocAsync; keep freed blocks in the device pool so
// that repeated calls do not pay for fresh allocations every time.
int keep_pool(int device)
{
    static std::mutex mu;
    static bool done[64] = {false};
    std::lock_guard<std::mutex> lk(mu);
    if (device < 0 || device >= 64 || done[device]) return 0;
    cudaMemPool_t pool;
    RADIAN_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    unsigned long long thr = ~0ull;
    RADIAN_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    done[device] = true;
    return 0;
}

// gate bit of context i: entropy(lm[context]) < r_threshold (decode.py:93, strict)
__global__ void gate_kernel(const double *__restrict__ entropy, size_t rows, double thr, uint32_t *__restrict__ gate)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < ((rows + 31) & ~(size_t)31); i += stride) {
        const bool g = i < rows && entropy[i] < thr;
        const unsigned b = __ballot_sync(0xffffffffu, g);
        if ((threadIdx.x & 31) == 0) gate[i >> 5] = b;
    }
}

int table_prepare_gate(radian_table *t, double r_threshold, cudaStream_t stream)
{
    if (t->gate_valid && t->gate_threshold == r_threshold) return 0;
    int dev = 0;
    RADIAN_CUDA(cudaGetDevice(&dev));
    DeviceInfo di;
    int rc = device_info(dev, &di);
    if (rc) return rc;
    gate_kernel<<<di.sm_count * 8, 256, 0, stream>>>(t->d_entropy, t->rows, r_threshold, t->d_gate);
    RADIAN_CUDA(cudaGetLastError());
    t->gate_threshold = r_threshold;
    t->gate_valid = 1;
    return 0;
}

}  // namespace radian

using namespace radian;

extern "C" const char *radian_last_error(void) { return g_err; }
extern "C" const char *radian_version(void) { return "radian_b200 0.1 (sm_100a)"; }

extern "C" int radian_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// entropy(np.asarray(row)) of decode.py:73-76: -sum(p*log(p) for p > 0), float64, left to right
static inline double row_entropy_host(const double *r)
{
    double s = 0.0;
    for (int i = 0; i < 4; ++i)
        if (r[i] > 0.0) s = s + r[i] * log(r[i]);
    return -s;
}

extern "C" int radian_table_create(const double *probs, int L, int device, radian_table_t **out)
{
    if (!probs || !out || L < 1 || L > RADIAN_MAX_CONTEXT) {
        set_error("radian_table_create: need probs, out and 1 <= L <= %d (got L=%d)", RADIAN_MAX_CONTEXT, L);
        return RADIAN_E_ARG;
    }
    if (radian_device_count() <= device || device < 0) {
        set_error("radian_table_create: CUDA device %d not available (no CPU fallback exists)", device);
        return RADIAN_E_CUDA;
    }
    RADIAN_CUDA(cudaSetDevice(device));
    const size_t rows = (size_t)1 << (2 * L);
    std::vector<double> ent(rows);
    unsigned nt = std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > 32) nt = 32;
    if (rows < (1u << 16)) nt = 1;
    std::vector<std::thread> th;
    for (unsigned k = 0; k < nt; ++k)
        th.emplace_back([&, k]() {
            const size_t lo = rows * k / nt, hi = rows * (k + 1) / nt;
            for (size_t i = lo; i < hi; ++i) ent[i] = row_entropy_host(probs + i * 4);
        });
    for (auto &x : th) x.join();

    radian_table *t = new radian_table();
    memset(t, 0, sizeof(*t));
    t->L = L;
    t->device = device;
    t->rows = rows;
    cudaError_t e;
    if ((e = cudaMalloc(&t->d_rows, rows * 4 * sizeof(double))) != cudaSuccess ||
        (e = cudaMalloc(&t->d_entropy, rows * sizeof(double))) != cudaSuccess ||
        (e = cudaMalloc(&t->d_gate, ((rows + 31) / 32) * sizeof(uint32_t))) != cudaSuccess ||
        (e = cudaMemcpy(t->d_rows, probs, rows * 4 * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(t->d_entropy, ent.data(), rows * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) {
        radian_table_destroy(t);
        return cuda_fail(e, "radian_table_create");
    }
    *out = t;
    return RADIAN_OK;
}

extern "C" int radian_table_destroy(radian_table_t *t)
{
    if (!t) return RADIAN_OK;
    cudaSetDevice(t->device);
    if (t->d_rows) cudaFree(t->d_rows);
    if (t->d_entropy) cudaFree(t->d_entropy);
    if (t->d_gate) cudaFree(t->d_gate);
    delete t;
    return RADIAN_OK;
}

extern "C" int radian_table_context_len(const radian_table_t *t) { return t ? t->L : 0; }

extern "C" int radian_table_entropies(const radian_table_t *t, double *out_host)
{
    if (!t || !out_host) {
        set_error("radian_table_entropies: null argument");
        return RADIAN_E_ARG;
    }
    RADIAN_CUDA(cudaSetDevice(t->device));
    RADIAN_CUDA(cudaMemcpy(out_host, t->d_entropy, t->rows * sizeof(double), cudaMemcpyDeviceToHost));
    return RADIAN_OK;
}

extern "C" size_t radian_decode_workspace_bytes(int device, int beam_width, int n_reads, int64_t max_frames,
                                                int64_t arena_nodes)
{
    if (beam_width < 1 || beam_width > RADIAN_MAX_BEAM_WIDTH || n_reads < 0) return 0;
    int slots = decode_max_slots(device, beam_width);
    if (slots <= 0) return 0;
    // one arena per resident read group; a launch never uses more groups than reads (rounded up to
    // whole CTAs of at most 16 groups)
    const int64_t need = ((int64_t)n_reads + 15) / 16 * 16;
    if (need < slots) slots = (int)(need < 16 ? 16 : need);
    const size_t cap = (size_t)decode_arena_cap(beam_width, max_frames, arena_nodes);
    return 256 + (size_t)slots * (cap + (size_t)decode_nursery()) * sizeof(uint32_t);
}

static int check_decode_args(const void *post, const int64_t *fo, int n_reads, int bw, const radian_table_t *table,
                             int L, uint8_t *out_seq, const int64_t *so, int64_t *out_len, double *out_score,
                             int32_t *out_status)
{
    if (n_reads < 0 || !fo || !so || !out_len || !out_score || !out_status || (n_reads > 0 && !out_seq)) {
        set_error("radian_decode_batch: null argument");
        return RADIAN_E_ARG;
    }
    (void)post;
    if (bw < 1 || bw > RADIAN_MAX_BEAM_WIDTH) {
        set_error("radian_decode_batch: beam_width %d outside 1..%d", bw, RADIAN_MAX_BEAM_WIDTH);
        return RADIAN_E_ARG;
    }
    if (table && L != table->L) {
        set_error("radian_decode_batch: len_context %d but the table holds %d-symbol contexts "
                  "(the reference raises KeyError at decode.py:83)", L, table->L);
        return RADIAN_E_CONTEXT;
    }
    return RADIAN_OK;
}

extern "C" int radian_decode_batch_dev(const void *post, int post_is_f64, const int64_t *frame_offsets, int n_reads,
                                       const int32_t *order, int64_t max_frames, int beam_width,
                                       const radian_table_t *table, int len_context, double s_threshold,
                                       double r_threshold, uint8_t *out_seq, const int64_t *seq_offsets,
                                       int64_t *out_len, double *out_score, int32_t *out_status,
                                       uint64_t *out_counters, int64_t arena_nodes, void *workspace,
                                       size_t workspace_bytes, radian_stream_t stream)
{
    int rc = check_decode_args(post, frame_offsets, n_reads, beam_width, table, len_context, out_seq, seq_offsets,
                               out_len, out_score, out_status);
    if (rc) return rc;
    if (n_reads == 0) return RADIAN_OK;
    int device = 0;
    RADIAN_CUDA(cudaGetDevice(&device));
    if (table && table->device != device) {
        set_error("radian_decode_batch_dev: table lives on device %d, current device is %d", table->device, device);
        return RADIAN_E_ARG;
    }
    const size_t need = radian_decode_workspace_bytes(device, beam_width, n_reads, max_frames, arena_nodes);
    const int64_t cap = decode_arena_cap(beam_width, max_frames, arena_nodes);
    if (cap >= (1ll << 29)) {
        set_error("radian_decode_batch_dev: arena of %lld nodes exceeds the 2^29 node limit", (long long)cap);
        return RADIAN_E_ARG;
    }
    if (!workspace || workspace_bytes < need) {
        set_error("radian_decode_batch_dev: workspace of %zu bytes needed, %zu given", need, workspace_bytes);
        return RADIAN_E_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (table) {
        rc = table_prepare_gate(const_cast<radian_table_t *>(table), r_threshold, st);
        if (rc) return rc;
    }
    RADIAN_CUDA(cudaMemsetAsync(workspace, 0, 256, st));
    DecodeArgs a;
    memset(&a, 0, sizeof(a));
    a.post = post;
    a.frame_offsets = frame_offsets;
    a.order = order;
    a.n_reads = n_reads;
    a.beam_width = beam_width;
    a.table = table ? table->d_rows : nullptr;
    a.gate = table ? table->d_gate : nullptr;
    a.L = table ? table->L : 0;
    a.s_thr = s_threshold;
    a.out_seq = out_seq;
    a.seq_offsets = seq_offsets;
    a.out_len = out_len;
    a.out_score = out_score;
    a.out_status = out_status;
    a.out_counters = (unsigned long long *)out_counters;
    a.queue = (int *)workspace;
    a.arena = (uint32_t *)((char *)workspace + 256);
    a.arena_cap = (int)cap;
    return decode_launch(a, post_is_f64 != 0, device, st);
}

// One pass over the reads listed in `sel` (indices into the caller's batch); results are written
// to the caller's arrays at those indices.  arena_nodes = 0 uses the default arena size.
static int decode_host_pass(const void *post, int post_is_f64, const int64_t *frame_offsets,
                            const std::vector<int32_t> &sel, int beam_width, const radian_table_t *table,
                            int len_context, double s_threshold, double r_threshold, uint8_t *out_seq,
                            const int64_t *seq_offsets, int64_t *out_len, double *out_score, int32_t *out_status,
                            uint64_t *out_counters, int64_t arena_nodes, int device)
{
    const int n = (int)sel.size();
    const size_t esz = post_is_f64 ? 8 : 4;
    // compact sub-batch: offsets relative to the packed copies
    std::vector<int64_t> fo(n + 1, 0), so(n + 1, 0);
    int64_t max_frames = 0;
    for (int i = 0; i < n; ++i) {
        const int r = sel[i];
        const int64_t T = frame_offsets[r + 1] - frame_offsets[r];
        fo[i + 1] = fo[i] + T;
        so[i + 1] = so[i] + (seq_offsets[r + 1] - seq_offsets[r]);
        max_frames = T > max_frames ? T : max_frames;
    }
    // longest reads first: the tail of the device work queue is then made of short reads
    std::vector<int32_t> order(n);
    for (int i = 0; i < n; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(),
                     [&](int32_t x, int32_t y) { return fo[x + 1] - fo[x] > fo[y + 1] - fo[y]; });
    const int64_t frames = fo[n], seq_bytes = so[n];
    const size_t ws_bytes = radian_decode_workspace_bytes(device, beam_width, n, max_frames, arena_nodes);
    {
        int krc = keep_pool(device);
        if (krc) return krc;
    }
    cudaStream_t st;
    RADIAN_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    void *d_post = nullptr, *d_ws = nullptr;
    int64_t *d_fo = nullptr, *d_so = nullptr, *d_len = nullptr;
    int32_t *d_order = nullptr, *d_status = nullptr;
    uint8_t *d_seq = nullptr;
    double *d_score = nullptr;
    uint64_t *d_cnt = nullptr;
    std::vector<uint8_t> h_seq((size_t)(seq_bytes ? seq_bytes : 1));
    std::vector<int64_t> h_len(n);
    std::vector<double> h_score((size_t)n * 2);
    std::vector<int32_t> h_status(n);
    std::vector<uint64_t> h_cnt((size_t)n * 2);
    int ret = RADIAN_OK;
    cudaError_t e;
#define TRY(x)                                   \
    if (ret == RADIAN_OK && (e = (x)) != cudaSuccess) ret = cuda_fail(e, #x)
    TRY(cudaMallocAsync(&d_post, (size_t)(frames ? frames : 1) * 5 * esz, st));
    TRY(cudaMallocAsync(&d_ws, ws_bytes, st));
    TRY(cudaMallocAsync(&d_fo, (size_t)(n + 1) * 8, st));
    TRY(cudaMallocAsync(&d_so, (size_t)(n + 1) * 8, st));
    TRY(cudaMallocAsync(&d_len, (size_t)n * 8, st));
    TRY(cudaMallocAsync(&d_order, (size_t)n * 4, st));
    TRY(cudaMallocAsync(&d_status, (size_t)n * 4, st));
    TRY(cudaMallocAsync(&d_seq, (size_t)(seq_bytes ? seq_bytes : 1), st));
    TRY(cudaMallocAsync(&d_score, (size_t)n * 16, st));
    if (out_counters) TRY(cudaMallocAsync(&d_cnt, (size_t)n * 16, st));
    // contiguous runs of selected reads are copied with one transfer each
    for (int i = 0; i < n && ret == RADIAN_OK;) {
        int j = i;
        while (j + 1 < n && sel[j + 1] == sel[j] + 1) ++j;
        const int64_t f0 = frame_offsets[sel[i]], f1 = frame_offsets[sel[j] + 1];
        if (f1 > f0)
            TRY(cudaMemcpyAsync((char *)d_post + (size_t)fo[i] * 5 * esz, (const char *)post + (size_t)f0 * 5 * esz,
                                (size_t)(f1 - f0) * 5 * esz, cudaMemcpyHostToDevice, st));
        i = j + 1;
    }
    TRY(cudaMemcpyAsync(d_fo, fo.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
    TRY(cudaMemcpyAsync(d_so, so.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
    TRY(cudaMemcpyAsync(d_order, order.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
    if (ret == RADIAN_OK)
        ret = radian_decode_batch_dev(d_post, post_is_f64, d_fo, n, d_order, max_frames, beam_width, table,
                                      len_context, s_threshold, r_threshold, d_seq, d_so, d_len, d_score, d_status,
                                      d_cnt, arena_nodes, d_ws, ws_bytes, st);
    TRY(cudaMemcpyAsync(h_seq.data(), d_seq, (size_t)seq_bytes, cudaMemcpyDeviceToHost, st));
    TRY(cudaMemcpyAsync(h_len.data(), d_len, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    TRY(cudaMemcpyAsync(h_score.data(), d_score, (size_t)n * 16, cudaMemcpyDeviceToHost, st));
    TRY(cudaMemcpyAsync(h_status.data(), d_status, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    if (out_counters) TRY(cudaMemcpyAsync(h_cnt.data(), d_cnt, (size_t)n * 16, cudaMemcpyDeviceToHost, st));
    TRY(cudaStreamSynchronize(st));
#undef TRY
    void *frees[] = {d_post, d_ws, d_fo, d_so, d_len, d_order, d_status, d_seq, d_score, d_cnt};
    for (void *p : frees)
        if (p) cudaFreeAsync(p, st);
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    if (ret != RADIAN_OK) return ret;
    for (int i = 0; i < n; ++i) {
        const int r = sel[i];
        out_len[r] = h_len[i];
        out_score[2 * r] = h_score[2 * i];
        out_score[2 * r + 1] = h_score[2 * i + 1];
        out_status[r] = h_status[i];
        if (out_counters) {
            out_counters[2 * r] = h_cnt[2 * i];
            out_counters[2 * r + 1] = h_cnt[2 * i + 1];
        }
        const int64_t slot = so[i + 1] - so[i];
        const int64_t ncopy = h_len[i] < slot ? h_len[i] : slot;
        if (ncopy > 0) memcpy(out_seq + seq_offsets[r], h_seq.data() + so[i], (size_t)ncopy);
    }
    return RADIAN_OK;
}

extern "C" int radian_decode_batch_host(const void *post, int post_is_f64, const int64_t *frame_offsets, int n_reads,
                                        int beam_width, const radian_table_t *table, int len_context,
                                        double s_threshold, double r_threshold, uint8_t *out_seq,
                                        const int64_t *seq_offsets, int64_t *out_len, double *out_score,
                                        int32_t *out_status, uint64_t *out_counters, int device)
{
    int rc = check_decode_args(post, frame_offsets, n_reads, beam_width, table, len_context, out_seq, seq_offsets,
                               out_len, out_score, out_status);
    if (rc) return rc;
    if (n_reads == 0) return RADIAN_OK;
    if (radian_device_count() <= device || device < 0) {
        set_error("radian_decode_batch_host: CUDA device %d not available (no CPU fallback exists)", device);
        return RADIAN_E_CUDA;
    }
    RADIAN_CUDA(cudaSetDevice(device));
    std::vector<int32_t> sel(n_reads);
    for (int i = 0; i < n_reads; ++i) {
        sel[i] = i;
        if (frame_offsets[i + 1] < frame_offsets[i] || seq_offsets[i + 1] < seq_offsets[i]) {
            set_error("radian_decode_batch_host: offsets not monotone at read %d", i);
            return RADIAN_E_ARG;
        }
    }
    rc = decode_host_pass(post, post_is_f64, frame_offsets, sel, beam_width, table, len_context, s_threshold,
                          r_threshold, out_seq, seq_offsets, out_len, out_score, out_status, out_counters, 0, device);
    if (rc) return rc;
    // reads whose labelings outgrew the default arena: once more with the exact worst case
    std::vector<int32_t> again;
    int64_t worst = 0;
    for (int i = 0; i < n_reads; ++i)
        if (out_status[i] == RADIAN_READ_TRIE_OVERFLOW) {
            again.push_back(i);
            const int64_t T = frame_offsets[i + 1] - frame_offsets[i];
            worst = T > worst ? T : worst;
        }
    if (!again.empty()) {
        rc = decode_host_pass(post, post_is_f64, frame_offsets, again, beam_width, table, len_context, s_threshold,
                              r_threshold, out_seq, seq_offsets, out_len, out_score, out_status, out_counters,
                              32 * (worst + 1) + 64, device);
        if (rc) return rc;
    }
    for (int i = 0; i < n_reads; ++i)
        if (out_status[i] != RADIAN_READ_OK) {
            set_error("radian_decode_batch_host: read %d failed with status %d", i, out_status[i]);
            return RADIAN_E_READ;
        }
    return RADIAN_OK;
}
