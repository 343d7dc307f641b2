// C ABI of libradian_b200.so: error reporting, the resident RNA table, decode entry points.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
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
        RADIAN_CUDA(cudaDeviceGetAttribute(&cache[device].smem_per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device));
        have[device] = true;
    }
    *out = cache[device];
    return 0;
}

// The _host entry points allocate stream-ordered from a pool of their own (one per device), which
// keeps freed blocks between calls so that repeated calls do not pay for fresh allocations.  The
// device's default pool -- shared with every other cudaMallocAsync user of the process -- is never
// touched: neither its release threshold nor its contents (radian_trim_memory trims only this pool).
static cudaMemPool_t g_pool[64] = {nullptr};

int keep_pool(int device, cudaMemPool_t *pool)
{
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (device < 0 || device >= 64) {
        set_error("bad device index %d", device);
        return RADIAN_E_ARG;
    }
    if (!g_pool[device]) {
        cudaMemPoolProps props;
        memset(&props, 0, sizeof(props));
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        cudaMemPool_t p = nullptr;
        RADIAN_CUDA(cudaMemPoolCreate(&p, &props));
        unsigned long long thr = ~0ull;  // keep freed blocks until radian_trim_memory
        RADIAN_CUDA(cudaMemPoolSetAttribute(p, cudaMemPoolAttrReleaseThreshold, &thr));
        g_pool[device] = p;
    }
    *pool = g_pool[device];
    return 0;
}

std::mutex &host_mutex(int device)
{
    static std::mutex mu[64];
    return mu[device >= 0 && device < 64 ? device : 0];
}

// gate bit of context i: entropy(lm[context]) < r_threshold (decode.py:93, strict)
// (a context that is absent from the model never opens the gate; a read that reaches it fails)
__global__ void gate_kernel(const double *__restrict__ entropy, const uint32_t *__restrict__ miss, size_t rows, double thr,
                            uint32_t *__restrict__ gate)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < ((rows + 31) & ~(size_t)31); i += stride) {
        const bool g = i < rows && entropy[i] < thr;
        unsigned b = __ballot_sync(0xffffffffu, g);
        if ((threadIdx.x & 31) == 0) {
            if (miss) b &= ~miss[i >> 5];
            gate[i >> 5] = b;
        }
    }
}

constexpr size_t kMaxGateMasks = 8;

int table_get_gate(const radian_table *t, double r_threshold, cudaStream_t stream, const uint32_t **d_bits)
{
    std::lock_guard<std::mutex> lk(t->mu);
    for (GateMask &g : t->gates)
        if (g.threshold == r_threshold) {
            g.last_use = ++t->tick;
            RADIAN_CUDA(cudaStreamWaitEvent(stream, g.ready, 0));
            *d_bits = g.d_bits;
            return 0;
        }
    if (t->gates.size() >= kMaxGateMasks) {
        // drop the mask that was used longest ago; a launch on some other stream may still be reading
        // it, so the device is drained first (a ninth distinct threshold on one table is rare)
        size_t lru = 0;
        for (size_t i = 1; i < t->gates.size(); ++i)
            if (t->gates[i].last_use < t->gates[lru].last_use) lru = i;
        RADIAN_CUDA(cudaDeviceSynchronize());
        cudaFree(t->gates[lru].d_bits);
        cudaEventDestroy(t->gates[lru].ready);
        t->gates.erase(t->gates.begin() + (long)lru);
    }
    DeviceInfo di;
    int rc = device_info(t->device, &di);
    if (rc) return rc;
    GateMask g;
    g.threshold = r_threshold;
    g.last_use = ++t->tick;
    g.d_bits = nullptr;
    g.ready = nullptr;
    RADIAN_CUDA(cudaMalloc(&g.d_bits, ((t->rows + 31) / 32) * sizeof(uint32_t)));
    cudaError_t e = cudaEventCreateWithFlags(&g.ready, cudaEventDisableTiming);
    if (e == cudaSuccess) {
        gate_kernel<<<di.sm_count * 8, 256, 0, stream>>>(t->d_entropy, t->d_miss, t->rows, r_threshold, g.d_bits);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaEventRecord(g.ready, stream);
    if (e != cudaSuccess) {
        cudaFree(g.d_bits);
        if (g.ready) cudaEventDestroy(g.ready);
        return cuda_fail(e, "table_get_gate");
    }
    t->gates.push_back(g);
    *d_bits = g.d_bits;
    return 0;
}

}  // namespace radian

using namespace radian;

extern "C" const char *radian_last_error(void) { return g_err; }
extern "C" const char *radian_version(void) { return "radian_b200 0.1 (sm_100a)"; }

// The _host entry points keep their device buffers in the stream-ordered pool between calls
// (keep_pool); this hands them back to the driver.
extern "C" int radian_trim_memory(int device)
{
    if (radian_device_count() <= device || device < 0) {
        set_error("radian_trim_memory: CUDA device %d not available", device);
        return RADIAN_E_CUDA;
    }
    std::lock_guard<std::mutex> host_lock(host_mutex(device));
    RADIAN_CUDA(cudaSetDevice(device));
    // (the streams of finished _host calls are synchronised before they return; nothing of this
    // library is in flight on the pool while the host mutex is held)
    if (g_pool[device]) RADIAN_CUDA(cudaMemPoolTrimTo(g_pool[device], 0));
    return RADIAN_OK;
}

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
    return radian_table_create_sparse(probs, nullptr, L, device, out);
}

extern "C" int radian_table_create_sparse(const double *probs, const uint8_t *present, int L, int device,
                                          radian_table_t **out)
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
    std::vector<double> vmax(nt, 0.0);
    std::vector<int> vbad(nt, 0);
    for (unsigned k = 0; k < nt; ++k)
        th.emplace_back([&, k]() {
            const size_t lo = rows * k / nt, hi = rows * (k + 1) / nt;
            double m = 0.0;
            int bad = 0;
            for (size_t i = lo; i < hi; ++i) {
                if (present && !present[i]) {
                    ent[i] = INFINITY;  // never below a threshold
                    continue;
                }
                ent[i] = row_entropy_host(probs + i * 4);
                for (int c = 0; c < 4; ++c) {
                    const double x = probs[i * 4 + c];
                    bad |= !(x >= 0.0 && x <= 4.0);  // also catches NaN
                    m = x > m ? x : m;
                }
            }
            vmax[k] = m;
            vbad[k] = bad;
        });
    for (auto &x : th) x.join();
    double tmax = 0.0;
    for (unsigned k = 0; k < nt; ++k) {
        tmax = vmax[k] > tmax ? vmax[k] : tmax;
        if (vbad[k]) {
            set_error("radian_table_create: table entries must be probabilities (finite, 0 <= p <= 4)");
            return RADIAN_E_ARG;
        }
    }

    radian_table *t = new radian_table();
    t->L = L;
    t->device = device;
    t->rows = rows;
    {
        // high word of the largest entry: what the decode kernel assumes of a row that is still on
        // its way from HBM (decode.cu, quiet-frame bound)
        uint64_t bits;
        memcpy(&bits, &tmax, 8);
        t->rcap = (int)(bits >> 32);
    }
    cudaError_t e;
    if ((e = cudaMalloc(&t->d_rows, rows * 4 * sizeof(double))) != cudaSuccess ||
        (e = cudaMalloc(&t->d_entropy, rows * sizeof(double))) != cudaSuccess ||
        (e = cudaMemcpy(t->d_rows, probs, rows * 4 * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(t->d_entropy, ent.data(), rows * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) {
        radian_table_destroy(t);
        return cuda_fail(e, "radian_table_create");
    }
    if (present) {
        std::vector<uint32_t> miss((rows + 31) / 32, 0u);
        bool any = false;
        for (size_t i = 0; i < rows; ++i)
            if (!present[i]) {
                miss[i >> 5] |= 1u << (i & 31);
                any = true;
            }
        if (any && ((e = cudaMalloc(&t->d_miss, miss.size() * 4)) != cudaSuccess ||
                    (e = cudaMemcpy(t->d_miss, miss.data(), miss.size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess)) {
            radian_table_destroy(t);
            return cuda_fail(e, "radian_table_create_sparse");
        }
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
    if (t->d_miss) cudaFree(t->d_miss);
    for (GateMask &g : t->gates) {
        cudaFree(g.d_bits);
        cudaEventDestroy(g.ready);
    }
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
    // one arena per resident read group; a launch never uses more warps than reads (a warp may hold a
    // single, exclusive read: up to four groups per read), plus the rounding to whole CTAs
    const int64_t need = ((int64_t)n_reads + 4) * 4;
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

static int decode_batch_dev_impl(const void *post, int post_is_f64, const int64_t *frame_offsets, int n_reads,
                                 const int32_t *order, int64_t max_frames, int64_t total_frames, int beam_width,
                                 const radian_table_t *table, int len_context, double s_threshold,
                                 double r_threshold, uint8_t *out_seq, const int64_t *seq_offsets, int64_t *out_len,
                                 double *out_score, int32_t *out_status, uint64_t *out_counters, int64_t arena_nodes,
                                 void *workspace, size_t workspace_bytes, const int *ready, cudaStream_t st)
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
    const uint32_t *d_gate = nullptr;
    if (table) {
        rc = table_get_gate(table, r_threshold, st, &d_gate);
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
    a.max_frames = max_frames;
    a.total_frames = total_frames;
    a.table = table ? table->d_rows : nullptr;
    a.gate = d_gate;
    a.miss = table ? table->d_miss : nullptr;
    a.L = table ? table->L : 0;
    a.rcap = table ? table->rcap : 0;
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
    a.ready = ready;
    return decode_launch(a, post_is_f64 != 0, device, st);
}

extern "C" int radian_decode_batch_dev(const void *post, int post_is_f64, const int64_t *frame_offsets, int n_reads,
                                       const int32_t *order, int64_t max_frames, int64_t total_frames, int beam_width,
                                       const radian_table_t *table, int len_context, double s_threshold,
                                       double r_threshold, uint8_t *out_seq, const int64_t *seq_offsets,
                                       int64_t *out_len, double *out_score, int32_t *out_status,
                                       uint64_t *out_counters, int64_t arena_nodes, void *workspace,
                                       size_t workspace_bytes, radian_stream_t stream)
{
    return decode_batch_dev_impl(post, post_is_f64, frame_offsets, n_reads, order, max_frames, total_frames, beam_width, table,
                                 len_context, s_threshold, r_threshold, out_seq, seq_offsets, out_len, out_score,
                                 out_status, out_counters, arena_nodes, workspace, workspace_bytes, nullptr,
                                 (cudaStream_t)stream);
}

// Page-locked scratch of the calling host thread (flag values going up, results coming down);
// grows, never shrinks, and is not returned at thread exit (the CUDA context may be gone by then).
static void *pinned_scratch(size_t bytes)
{
    static thread_local void *p = nullptr;
    static thread_local size_t cap = 0;
    if (bytes <= cap) return p;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    const size_t want = bytes + bytes / 4 + 4096;
    if (cudaHostAlloc(&p, want, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        p = nullptr;
        return nullptr;
    }
    cap = want;
    return p;
}


// Streams and events of the calling host thread's decode calls, made once per (thread, device) and
// kept: the reference's own calling pattern is one read per call (basecall.py:70-123), which should
// not pay for five creations and destructions every time.
struct HostStreams {
    int device = -1;
    cudaStream_t st = nullptr, cs[2] = {nullptr, nullptr};
    cudaEvent_t ev = nullptr, ev_last = nullptr;
};

static HostStreams *host_streams(int device)
{
    static thread_local HostStreams hs[4];
    HostStreams *h = nullptr;
    for (auto &x : hs)
        if (x.device == device) return &x;
    for (auto &x : hs)
        if (x.device < 0) {
            h = &x;
            break;
        }
    if (!h) {  // more than four devices driven by one thread: recycle the first slot
        h = &hs[0];
        cudaStreamDestroy(h->st);
        cudaStreamDestroy(h->cs[0]);
        cudaStreamDestroy(h->cs[1]);
        cudaEventDestroy(h->ev);
        cudaEventDestroy(h->ev_last);
        *h = HostStreams();
    }
    if (cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->cs[0], cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&h->cs[1], cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->ev_last, cudaEventDisableTiming) != cudaSuccess) {
        cuda_fail(cudaGetLastError(), "host_streams");
        return nullptr;
    }
    h->device = device;
    return h;
}

// ---- staged upload -----------------------------------------------------------------------------
// Pageable host memory reaches the device at ~10 GB/s through cudaMemcpyAsync (the driver stages
// it on the calling thread), and page-locked memory at 80 % of the link when every read is its own
// transfer.  Here host threads gather the reads, in queue order, into a ring of page-locked slots
// and the calling thread sends every filled slot with one large copy: the source may be any
// memory, and the copy engine only sees 16 MB transfers.
// (one host thread fills one slot at a time at 6-8 GB/s: the number of slots is the number of threads
// that can work ahead of the copy engine, and it takes about eight of them to outrun the link)
constexpr int kStageSlots = 12;
constexpr size_t kStageBytes = (size_t)16 << 20;

struct StageRing {
    char *slot[kStageSlots] = {};
    cudaEvent_t ev[kStageSlots] = {};
    int *flags = nullptr;  // page-locked flag values
    size_t bytes = 0, n_flags = 0;
};

static StageRing *stage_ring(size_t slot_bytes, size_t n_flags)
{
    static thread_local StageRing ring;
    if (slot_bytes > ring.bytes) {
        for (int k = 0; k < kStageSlots; ++k) {
            if (ring.slot[k]) cudaFreeHost(ring.slot[k]);
            ring.slot[k] = nullptr;
            if (cudaHostAlloc((void **)&ring.slot[k], slot_bytes, cudaHostAllocDefault) != cudaSuccess) {
                cudaGetLastError();
                ring.bytes = 0;
                return nullptr;
            }
            if (!ring.ev[k] && cudaEventCreateWithFlags(&ring.ev[k], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        }
        ring.bytes = slot_bytes;
    }
    if (n_flags > ring.n_flags) {
        if (ring.flags) cudaFreeHost(ring.flags);
        ring.flags = nullptr;
        if (cudaHostAlloc((void **)&ring.flags, (n_flags + 64) * sizeof(int), cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            ring.n_flags = 0;
            return nullptr;
        }
        ring.n_flags = n_flags + 64;
    }
    return &ring;
}

struct StageSeg {
    int k0, k1;
};

struct StagePlan {
    std::vector<StageSeg> segs;
    StageRing *ring = nullptr;
};

// Segments of the queue that fit a slot, and the ring itself.  Everything that allocates happens
// here, BEFORE the kernel is launched: a page-locked allocation can wait for the device to go
// idle, which a kernel waiting for its input never does.
static int stage_plan(const std::vector<int64_t> &fo, size_t row, StagePlan *plan)
{
    const int n = (int)fo.size() - 1;
    size_t biggest = 0;
    for (int k = 0; k < n; ++k) biggest = std::max(biggest, (size_t)(fo[k + 1] - fo[k]) * row);
    const size_t slot_bytes = std::max(kStageBytes, biggest);
    for (int k = 0; k < n;) {
        int j = k;
        size_t bytes = (size_t)(fo[k + 1] - fo[k]) * row;
        while (j + 1 < n && bytes + (size_t)(fo[j + 2] - fo[j + 1]) * row <= slot_bytes) {
            ++j;
            bytes += (size_t)(fo[j + 1] - fo[j]) * row;
        }
        plan->segs.push_back({k, j + 1});
        k = j + 1;
    }
    plan->ring = stage_ring(slot_bytes, plan->segs.size());
    if (!plan->ring) {
        set_error("radian_decode_batch_host: cannot page-lock the staging ring (%zu bytes per slot)", slot_bytes);
        return RADIAN_E_CUDA;
    }
    return RADIAN_OK;
}

// reads in queue order k = 0..n-1: source src[k], device frames fo[k]..fo[k+1]
static int staged_upload(size_t row, const std::vector<const char *> &src,
                         const std::vector<int64_t> &fo, const StagePlan &plan, char *d_post, int *d_ready,
                         cudaStream_t cs[2], cudaEvent_t ev_last, int device)
{
    const std::vector<StageSeg> &segs = plan.segs;
    StageRing *ring = plan.ring;
    const int nseg = (int)segs.size();
    std::vector<std::atomic<int>> filled((size_t)nseg), issued((size_t)nseg);
    for (int i = 0; i < nseg; ++i) filled[i].store(0), issued[i].store(0);
    std::atomic<int> next(0), failed(0);
    unsigned nw = std::thread::hardware_concurrency();
    nw = nw > 4 ? nw - 2 : 2;  // leave room for the submitting thread
    nw = nw > 12 ? 12 : nw;
    if ((int)nw > nseg) nw = (unsigned)nseg;
    std::vector<std::thread> workers;
    for (unsigned w = 0; w < nw; ++w)
        workers.emplace_back([&]() {
            cudaSetDevice(device);
            for (;;) {
                const int sgm = next.fetch_add(1);
                if (sgm >= nseg || failed.load()) break;
                const int slot = sgm % kStageSlots;
                if (sgm >= kStageSlots) {  // the slot's previous transfer must have left it
                    while (!issued[sgm - kStageSlots].load(std::memory_order_acquire) && !failed.load()) std::this_thread::yield();
                    if (cudaEventSynchronize(ring->ev[slot]) != cudaSuccess) failed.store(1);
                }
                char *dst = ring->slot[slot];
                for (int k = segs[sgm].k0; k < segs[sgm].k1; ++k) {
                    const size_t bytes = (size_t)(fo[k + 1] - fo[k]) * row;
                    memcpy(dst, src[k], bytes);
                    dst += bytes;
                }
                filled[sgm].store(1, std::memory_order_release);
            }
        });
    int ret = RADIAN_OK;
    cudaError_t e = cudaSuccess;
    const char *stall_env = getenv("RADIAN_TEST_STALL_MS");  // test hook: hold the copies back midway
    for (int sgm = 0; sgm < nseg && ret == RADIAN_OK; ++sgm) {
        if (stall_env && sgm == nseg / 2) {
            cudaStreamSynchronize(cs[0]);
            cudaStreamSynchronize(cs[1]);
            std::this_thread::sleep_for(std::chrono::milliseconds(atoi(stall_env)));
        }
        while (!filled[sgm].load(std::memory_order_acquire) && !failed.load()) std::this_thread::yield();
        if (failed.load()) {
            set_error("radian_decode_batch_host: staging worker failed");
            ret = RADIAN_E_CUDA;
            break;
        }
        const int slot = sgm % kStageSlots, which = sgm & 1;
        const size_t bytes = (size_t)(fo[segs[sgm].k1] - fo[segs[sgm].k0]) * row;
        ring->flags[sgm] = segs[sgm].k1;
        if (bytes && (e = cudaMemcpyAsync(d_post + (size_t)fo[segs[sgm].k0] * row, ring->slot[slot], bytes,
                                          cudaMemcpyHostToDevice, cs[which])) != cudaSuccess)
            ret = cuda_fail(e, "staged cudaMemcpyAsync");
        if (ret == RADIAN_OK && (e = cudaMemcpyAsync(d_ready + which, &ring->flags[sgm], 4, cudaMemcpyHostToDevice,
                                                     cs[which])) != cudaSuccess)
            ret = cuda_fail(e, "staged flag copy");
        if (ret == RADIAN_OK && (e = cudaEventRecord(ring->ev[slot], cs[which])) != cudaSuccess)
            ret = cuda_fail(e, "cudaEventRecord");
        issued[sgm].store(1, std::memory_order_release);
        if (ret == RADIAN_OK && sgm + 1 == nseg) {
            // the other stream's counter stops short of n: let it catch up once everything landed
            if ((e = cudaEventRecord(ev_last, cs[which])) != cudaSuccess ||
                (e = cudaStreamWaitEvent(cs[which ^ 1], ev_last, 0)) != cudaSuccess ||
                (e = cudaMemcpyAsync(d_ready + (which ^ 1), &ring->flags[sgm], 4, cudaMemcpyHostToDevice,
                                     cs[which ^ 1])) != cudaSuccess)
                ret = cuda_fail(e, "final flag copy");
        }
    }
    if (ret != RADIAN_OK) failed.store(1);
    for (auto &t : workers) t.join();
    return ret;
}

// One streamed pass over the reads listed in `sel` (indices into the caller's batch); results are
// written to the caller's arrays at those indices.  arena_nodes = 0 uses the default arena size.
//
// The kernel is launched first and the posteriors follow: the reads are queued longest first, the
// copy stream sends them over PCIe in exactly that order and publishes, every few megabytes, how
// many reads have landed (`ready`); a read group that pops a read which is still in flight waits
// for it.  The transfer, which is the slower of the two at ~20 B per frame, is therefore the only
// thing on the critical path; the decode of read k overlaps the copies of reads k+1...
static int decode_host_pass(const void *post, const void *const *read_ptrs, int post_is_f64, const int64_t *frame_offsets,
                            const std::vector<int32_t> &sel, int beam_width, const radian_table_t *table,
                            int len_context, double s_threshold, double r_threshold, uint8_t *out_seq,
                            const int64_t *seq_offsets, int64_t *out_len, double *out_score, int32_t *out_status,
                            uint64_t *out_counters, int64_t arena_nodes, int device)
{
    const int n = (int)sel.size();
    const size_t esz = post_is_f64 ? 8 : 4;
    const size_t row = 5 * esz;
    static const bool trace = getenv("RADIAN_TRACE") != nullptr;
    // Kernel-replay profilers (ncu) run one kernel at a time and hold back the copy streams, so a
    // kernel that waits for its input to arrive never finishes under them: RADIAN_HOST_COPY_FIRST
    // makes this call copy everything before it launches.
    // (also switched on when the process runs under an injected CUDA profiler)
    static const bool copy_first = getenv("RADIAN_HOST_COPY_FIRST") != nullptr ||
                                   getenv("CUDA_INJECTION64_PATH") != nullptr ||
                                   getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") != nullptr;
    double tr[8] = {0};
    auto stamp = [&](int i) {
        if (trace) tr[i] = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
    };
    stamp(0);
    constexpr size_t kPublishBytes = 16u << 20;  // a copy (and a flag update) after at most this much payload
    // Queue position k holds read sel[q[k]].  The queue order is also the order in which the reads
    // travel, and a read can only be started once it has landed; the link moves a frame ~750 times
    // faster than one read group decodes it, so a long read must not arrive late.  But the link also
    // wants large copies (a copy costs ~4 us before its first byte: 16 MB copies reach 98 % of the
    // link, one copy per read 80 %).  So the caller's buffer is cut into chunks of adjacent reads
    // (up to 16 MB, 1/64 of the batch for small ones) and the chunks travel in the order of their
    // longest read, longest first; inside a chunk the caller's order stands.
    std::vector<int32_t> q(n);
    auto T_of = [&](int i) { return frame_offsets[sel[i] + 1] - frame_offsets[sel[i]]; };
    // where read sel[i] is in host memory: one buffer with offsets, or one pointer per read
    auto src_of = [&](int i) -> const char * {
        return read_ptrs ? (const char *)read_ptrs[sel[i]] : (const char *)post + (size_t)frame_offsets[sel[i]] * row;
    };
    {
        int64_t total = 0;
        for (int i = 0; i < n; ++i) total += T_of(i);
        size_t chunk_bytes = (size_t)total * row / 64;
        chunk_bytes = chunk_bytes < ((size_t)256 << 10) ? ((size_t)256 << 10) : chunk_bytes > kPublishBytes ? kPublishBytes : chunk_bytes;
        struct Chunk {
            int first, last;  // reads [first, last) of sel
            int64_t longest;
        };
        std::vector<Chunk> chunks;
        for (int i = 0; i < n;) {
            int j = i + 1;
            size_t bytes = (size_t)T_of(i) * row;
            int64_t longest = T_of(i);
            while (j < n && src_of(j) == src_of(j - 1) + (size_t)T_of(j - 1) * row && bytes + (size_t)T_of(j) * row <= chunk_bytes) {
                bytes += (size_t)T_of(j) * row;
                longest = std::max(longest, T_of(j));
                ++j;
            }
            chunks.push_back({i, j, longest});
            i = j;
        }
        std::stable_sort(chunks.begin(), chunks.end(), [](const Chunk &x, const Chunk &y) { return x.longest > y.longest; });
        int k = 0;
        for (const Chunk &c : chunks)
            for (int i = c.first; i < c.last; ++i) q[k++] = i;
    }
    // the device sees the reads renumbered by queue position and packed in that order
    std::vector<int64_t> fo(n + 1, 0), so(n + 1, 0);
    int64_t max_frames = 0;
    for (int k = 0; k < n; ++k) {
        const int r = sel[q[k]];
        const int64_t T = T_of(q[k]);
        fo[k + 1] = fo[k] + T;
        so[k + 1] = so[k] + (seq_offsets[r + 1] - seq_offsets[r]);
        max_frames = T > max_frames ? T : max_frames;
    }
    const int64_t frames = fo[n], seq_bytes = so[n];
    const size_t ws_bytes = radian_decode_workspace_bytes(device, beam_width, n, max_frames, arena_nodes);
    cudaMemPool_t pool_ = nullptr;  // this library's own stream-ordered pool on the device
    {
        int krc = keep_pool(device, &pool_);
        if (krc) return krc;
    }
    // copy plan: runs of queue-adjacent reads that are also adjacent in the caller's buffer go in
    // one transfer; `pub` = reads published after the transfer
    struct Xfer {
        int64_t dst_frame;
        const char *src;
        int64_t n_frames;
        int pub;  // reads landed after this transfer, or 0 when no flag update follows it
    };
    std::vector<Xfer> plan;
    size_t unpublished = 0;
    for (int k = 0; k < n;) {
        int j = k;
        size_t bytes = (size_t)T_of(q[k]) * row;
        while (j + 1 < n && src_of(q[j + 1]) == src_of(q[j]) + (size_t)T_of(q[j]) * row && bytes < kPublishBytes) {
            ++j;
            bytes += (size_t)T_of(q[j]) * row;
        }
        unpublished += bytes;
        const bool pub = unpublished >= kPublishBytes || j + 1 == n;
        plan.push_back({fo[k], src_of(q[k]), fo[j + 1] - fo[k], pub ? j + 1 : 0});
        if (pub) unpublished = 0;
        k = j + 1;
    }
    // page-locked scratch: flag values | len | score | status | counters | sequences
    const size_t o_flag = 0;
    const size_t o_len = (plan.size() * 4 + 15) & ~(size_t)15;
    const size_t o_score = o_len + (size_t)n * 8;
    const size_t o_status = o_score + (size_t)n * 16;
    const size_t o_cnt = (o_status + (size_t)n * 4 + 15) & ~(size_t)15;
    const size_t o_seq = o_cnt + (out_counters ? (size_t)n * 32 : 0);
    char *hp = (char *)pinned_scratch(o_seq + (size_t)seq_bytes + 16);
    if (!hp) {
        set_error("radian_decode_batch_host: cannot page-lock %zu bytes of host scratch", o_seq + (size_t)seq_bytes);
        return RADIAN_E_CUDA;
    }
    int *h_flag = (int *)(hp + o_flag);
    int64_t *h_len = (int64_t *)(hp + o_len);
    double *h_score = (double *)(hp + o_score);
    int32_t *h_status = (int32_t *)(hp + o_status);
    uint64_t *h_cnt = (uint64_t *)(hp + o_cnt);
    uint8_t *h_seq = (uint8_t *)(hp + o_seq);
    for (size_t i = 0; i < plan.size(); ++i) h_flag[i] = plan[i].pub;

    // pageable sources (and RADIAN_HOST_STAGE=1) go through the staging ring, which is set up now
    bool staged = getenv("RADIAN_HOST_STAGE") != nullptr && getenv("RADIAN_HOST_STAGE")[0] != '0';
    if (!staged && getenv("RADIAN_HOST_STAGE") == nullptr && frames > 0) {
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, src_of(q[0])) != cudaSuccess) {
            cudaGetLastError();
            staged = true;
        } else {
            staged = pa.type == cudaMemoryTypeUnregistered;
        }
    }
    StagePlan splan;
    if (staged) {
        const int prc = stage_plan(fo, row, &splan);
        if (prc) return prc;
    }
    HostStreams *hs = host_streams(device);
    if (!hs) return RADIAN_E_CUDA;
    cudaStream_t st = hs->st, cs[2] = {hs->cs[0], hs->cs[1]};
    cudaEvent_t ev = hs->ev, ev_last = hs->ev_last;
    void *d_post = nullptr, *d_ws = nullptr;
    int64_t *d_fo = nullptr, *d_so = nullptr, *d_len = nullptr;
    int32_t *d_status = nullptr, *d_order = nullptr;
    int *d_ready = nullptr;
    uint8_t *d_seq = nullptr;
    double *d_score = nullptr;
    uint64_t *d_cnt = nullptr;
    int ret = RADIAN_OK;
    bool launched = false;
    cudaError_t e;
#define TRY(x)                                   \
    if (ret == RADIAN_OK && (e = (x)) != cudaSuccess) ret = cuda_fail(e, #x)
    TRY(cudaMallocFromPoolAsync(&d_post, (size_t)(frames ? frames : 1) * row, pool_, st));
    TRY(cudaMallocFromPoolAsync(&d_ws, ws_bytes, pool_, st));
    TRY(cudaMallocFromPoolAsync(&d_fo, (size_t)(n + 1) * 8, pool_, st));
    TRY(cudaMallocFromPoolAsync(&d_so, (size_t)(n + 1) * 8, pool_, st));
    TRY(cudaMallocFromPoolAsync(&d_len, (size_t)n * 8, pool_, st));
    TRY(cudaMallocFromPoolAsync(&d_status, (size_t)n * 4, pool_, st));
    TRY(cudaMallocFromPoolAsync(&d_ready, 256, pool_, st));
    TRY(cudaMallocFromPoolAsync(&d_order, (size_t)n * 4, pool_, st));
    TRY(cudaMallocFromPoolAsync(&d_seq, (size_t)(seq_bytes ? seq_bytes : 1), pool_, st));
    TRY(cudaMallocFromPoolAsync(&d_score, (size_t)n * 16, pool_, st));
    if (out_counters) TRY(cudaMallocFromPoolAsync(&d_cnt, (size_t)n * 32, pool_, st));
    TRY(cudaMemcpyAsync(d_fo, fo.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
    TRY(cudaMemcpyAsync(d_so, so.data(), (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, st));
    TRY(cudaMemsetAsync(d_ready, 0, 256, st));
    TRY(cudaMemsetAsync(d_status, 0xff, (size_t)n * 4, st));  // -1 = not run (stalled-transfer guard)
    TRY(cudaEventRecord(ev, st));
    // the copies need the allocations and the cleared counters
    TRY(cudaStreamWaitEvent(cs[0], ev, 0));
    TRY(cudaStreamWaitEvent(cs[1], ev, 0));
    if (ret == RADIAN_OK && !copy_first) {
        ret = decode_batch_dev_impl(d_post, post_is_f64, d_fo, n, nullptr, max_frames, frames, beam_width, table,
                                    len_context, s_threshold, r_threshold, d_seq, d_so, d_len, d_score, d_status,
                                    d_cnt, arena_nodes, d_ws, ws_bytes, d_ready, st);
        launched = (ret == RADIAN_OK);
    }
    stamp(1);
    // Segments (the transfers up to and including a publishing one) alternate between two copy
    // streams, so that the fixed gap between stream-ordered copies of one stream is covered by
    // the other stream's transfer; stream s publishes into d_ready[s].
    if (staged && ret == RADIAN_OK) {
        std::vector<const char *> src((size_t)n);
        for (int k = 0; k < n; ++k) src[k] = src_of(q[k]);
        ret = staged_upload(row, src, fo, splan, (char *)d_post, d_ready, cs, ev_last, device);
    }
    int which = 0;
    const char *stall_env = getenv("RADIAN_TEST_STALL_MS");  // test hook: hold the copies back midway
    for (size_t i = 0; i < plan.size() && ret == RADIAN_OK && !staged; ++i) {
        const Xfer &x = plan[i];
        if (stall_env && i == plan.size() / 2) {
            cudaStreamSynchronize(cs[0]);
            cudaStreamSynchronize(cs[1]);
            std::this_thread::sleep_for(std::chrono::milliseconds(atoi(stall_env)));
        }
        if (x.n_frames > 0)
            TRY(cudaMemcpyAsync((char *)d_post + (size_t)x.dst_frame * row,
                                x.src, (size_t)x.n_frames * row,
                                cudaMemcpyHostToDevice, cs[which]));
        if (x.pub) {
            TRY(cudaMemcpyAsync(d_ready + which, &h_flag[i], 4, cudaMemcpyHostToDevice, cs[which]));
            if (i + 1 == plan.size()) {
                // the other stream's counter stops short of n: let it catch up once everything landed
                TRY(cudaEventRecord(ev_last, cs[which]));
                TRY(cudaStreamWaitEvent(cs[which ^ 1], ev_last, 0));
                TRY(cudaMemcpyAsync(d_ready + (which ^ 1), &h_flag[i], 4, cudaMemcpyHostToDevice, cs[which ^ 1]));
            }
            which ^= 1;
        }
    }
    if (launched && ret != RADIAN_OK) {
        // a transfer failed while the kernel is waiting for it: release the waiters so that the
        // launch drains (its results are discarded)
        cudaMemsetAsync(d_ready, 0x7f, 8, cs[0]);
    }
    if (copy_first) {
        for (int k = 0; k < 2; ++k) {
            TRY(cudaEventRecord(ev_last, cs[k]));
            TRY(cudaStreamWaitEvent(st, ev_last, 0));
        }
        if (ret == RADIAN_OK)
            ret = decode_batch_dev_impl(d_post, post_is_f64, d_fo, n, nullptr, max_frames, frames, beam_width, table,
                                        len_context, s_threshold, r_threshold, d_seq, d_so, d_len, d_score,
                                        d_status, d_cnt, arena_nodes, d_ws, ws_bytes, nullptr, st);
    }
    stamp(2);
    if (launched) {
        // Reads the kernel gave up on because the transfer stood still (status still -1): once the
        // copies are through, a second launch decodes them from the resident data.
        TRY(cudaMemcpyAsync(h_status, d_status, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        TRY(cudaStreamSynchronize(cs[0]));
        TRY(cudaStreamSynchronize(cs[1]));
        TRY(cudaStreamSynchronize(st));
        std::vector<int32_t> redo;
        for (int k = 0; k < n && ret == RADIAN_OK; ++k)
            if (h_status[k] == -1) redo.push_back(k);
        if (!redo.empty() && ret == RADIAN_OK) {
            if (trace) fprintf(stderr, "[radian] host pass: transfer stalled, %zu reads decoded by a second launch\n", redo.size());
            TRY(cudaMemcpyAsync(d_order, redo.data(), redo.size() * 4, cudaMemcpyHostToDevice, st));
            if (ret == RADIAN_OK)
                ret = decode_batch_dev_impl(d_post, post_is_f64, d_fo, (int)redo.size(), d_order, max_frames, 0,
                                            beam_width, table, len_context, s_threshold, r_threshold, d_seq, d_so,
                                            d_len, d_score, d_status, d_cnt, arena_nodes, d_ws, ws_bytes, nullptr, st);
            TRY(cudaStreamSynchronize(st));  // redo must outlive the copy of the order array
        }
    }
    TRY(cudaMemcpyAsync(h_seq, d_seq, (size_t)seq_bytes, cudaMemcpyDeviceToHost, st));
    TRY(cudaMemcpyAsync(h_len, d_len, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    TRY(cudaMemcpyAsync(h_score, d_score, (size_t)n * 16, cudaMemcpyDeviceToHost, st));
    TRY(cudaMemcpyAsync(h_status, d_status, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    if (out_counters) TRY(cudaMemcpyAsync(h_cnt, d_cnt, (size_t)n * 32, cudaMemcpyDeviceToHost, st));
    TRY(cudaStreamSynchronize(cs[0]));
    TRY(cudaStreamSynchronize(cs[1]));
    stamp(3);
    TRY(cudaStreamSynchronize(st));
    stamp(4);
#undef TRY
    void *frees[] = {d_post, d_ws, d_fo, d_so, d_len, d_status, d_order, d_ready, d_seq, d_score, d_cnt};
    for (void *p : frees)
        if (p) cudaFreeAsync(p, st);
    cudaStreamSynchronize(st);
    if (ret != RADIAN_OK) return ret;
    // results back into the caller's order (a few host threads for a large batch: the symbols of
    // 16 384 reads are 25 MB in as many pieces)
    auto scatter = [&](int k0, int k1) {
        for (int k = k0; k < k1; ++k) {
            const int r = sel[q[k]];
            out_len[r] = h_len[k];
            out_score[2 * r] = h_score[2 * k];
            out_score[2 * r + 1] = h_score[2 * k + 1];
            out_status[r] = h_status[k];
            if (out_counters) {
                for (int c = 0; c < 4; ++c) out_counters[4 * r + c] = h_cnt[4 * k + c];
            }
            const int64_t slot = so[k + 1] - so[k];
            const int64_t ncopy = h_len[k] < slot ? h_len[k] : slot;
            if (ncopy > 0) memcpy(out_seq + seq_offsets[r], h_seq + so[k], (size_t)ncopy);
        }
    };
    if (n < 2048) {
        scatter(0, n);
    } else {
        unsigned nt = std::thread::hardware_concurrency();
        nt = nt < 2 ? 2 : nt > 8 ? 8 : nt;
        std::vector<std::thread> th;
        for (unsigned w = 0; w < nt; ++w) th.emplace_back(scatter, (int)((int64_t)n * w / nt), (int)((int64_t)n * (w + 1) / nt));
        for (auto &t : th) t.join();
    }
    stamp(5);
    if (trace)
        fprintf(stderr, "[radian] host pass: %d reads, %zu transfers, %.2f GB | plan+alloc+launch %.1f ms, submit %.1f ms, "
                "copies done +%.1f ms, kernel+results done +%.1f ms, scatter %.1f ms, total %.1f ms\n",
                n, plan.size(), (double)frames * row / 1e9, 1e3 * (tr[1] - tr[0]), 1e3 * (tr[2] - tr[1]),
                1e3 * (tr[3] - tr[2]), 1e3 * (tr[4] - tr[3]), 1e3 * (tr[5] - tr[4]), 1e3 * (tr[5] - tr[0]));
    return RADIAN_OK;
}

static int decode_batch_host_impl(const void *post, const void *const *read_ptrs, int post_is_f64,
                                  const int64_t *frame_offsets, int n_reads, int beam_width, const radian_table_t *table,
                                  int len_context, double s_threshold, double r_threshold, uint8_t *out_seq,
                                  const int64_t *seq_offsets, int64_t *out_len, double *out_score, int32_t *out_status,
                                  uint64_t *out_counters, int device)
{
    int rc = check_decode_args(post, frame_offsets, n_reads, beam_width, table, len_context, out_seq, seq_offsets,
                               out_len, out_score, out_status);
    if (rc) return rc;
    if (n_reads == 0) return RADIAN_OK;
    if (radian_device_count() <= device || device < 0) {
        set_error("radian_decode_batch_host: CUDA device %d not available (no CPU fallback exists)", device);
        return RADIAN_E_CUDA;
    }
    std::lock_guard<std::mutex> host_lock(host_mutex(device));
    RADIAN_CUDA(cudaSetDevice(device));
    std::vector<int32_t> sel(n_reads);
    for (int i = 0; i < n_reads; ++i) {
        sel[i] = i;
        if (frame_offsets[i + 1] < frame_offsets[i] || seq_offsets[i + 1] < seq_offsets[i]) {
            set_error("radian_decode_batch_host: offsets not monotone at read %d", i);
            return RADIAN_E_ARG;
        }
    }
    rc = decode_host_pass(post, read_ptrs, post_is_f64, frame_offsets, sel, beam_width, table, len_context, s_threshold,
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
        rc = decode_host_pass(post, read_ptrs, post_is_f64, frame_offsets, again, beam_width, table, len_context,
                              s_threshold, r_threshold, out_seq, seq_offsets, out_len, out_score, out_status,
                              out_counters, (int64_t)128 * (worst + 1) + 64, device);
        if (rc) return rc;
    }
    for (int i = 0; i < n_reads; ++i)
        if (out_status[i] != RADIAN_READ_OK) {
            set_error("radian_decode_batch_host: read %d failed with status %d", i, out_status[i]);
            return RADIAN_E_READ;
        }
    return RADIAN_OK;
}

extern "C" int radian_decode_batch_host(const void *post, int post_is_f64, const int64_t *frame_offsets, int n_reads,
                                        int beam_width, const radian_table_t *table, int len_context,
                                        double s_threshold, double r_threshold, uint8_t *out_seq,
                                        const int64_t *seq_offsets, int64_t *out_len, double *out_score,
                                        int32_t *out_status, uint64_t *out_counters, int device)
{
    return decode_batch_host_impl(post, nullptr, post_is_f64, frame_offsets, n_reads, beam_width, table, len_context,
                                  s_threshold, r_threshold, out_seq, seq_offsets, out_len, out_score, out_status,
                                  out_counters, device);
}

extern "C" int radian_decode_batch_host_reads(const void *const *reads, const int64_t *n_frames, int post_is_f64,
                                              int n_reads, int beam_width, const radian_table_t *table,
                                              int len_context, double s_threshold, double r_threshold,
                                              uint8_t *out_seq, const int64_t *seq_offsets, int64_t *out_len,
                                              double *out_score, int32_t *out_status, uint64_t *out_counters,
                                              int device)
{
    if (n_reads < 0 || (n_reads > 0 && (!reads || !n_frames))) {
        set_error("radian_decode_batch_host_reads: null argument");
        return RADIAN_E_ARG;
    }
    std::vector<int64_t> fo((size_t)n_reads + 1, 0);
    for (int i = 0; i < n_reads; ++i) {
        if (n_frames[i] < 0 || (n_frames[i] > 0 && !reads[i])) {
            set_error("radian_decode_batch_host_reads: read %d has no matrix", i);
            return RADIAN_E_ARG;
        }
        fo[i + 1] = fo[i] + n_frames[i];
    }
    return decode_batch_host_impl(nullptr, reads, post_is_f64, fo.data(), n_reads, beam_width, table, len_context,
                                  s_threshold, r_threshold, out_seq, seq_offsets, out_len, out_score, out_status,
                                  out_counters, device);
}

// FASTA text of a decoded batch, formed on the host by a few threads (basecall.py:129 writes
// f">{read.read_id}\n{sequence[::-1]}\n" per read: the decoder's output is in sequencing order,
// 3'->5', and is reversed here).
extern "C" int radian_fasta_records_host(const uint8_t *seq, const int64_t *seq_offsets, const int64_t *len,
                                         int n_reads, const char *ids, const int64_t *id_offsets, const char *bases,
                                         char *out, const int64_t *out_offsets)
{
    if (n_reads < 0 || (n_reads > 0 && (!seq || !seq_offsets || !len || !ids || !id_offsets || !bases || !out || !out_offsets))) {
        set_error("radian_fasta_records_host: null argument");
        return RADIAN_E_ARG;
    }
    for (int r = 0; r < n_reads; ++r) {
        const int64_t need = 1 + (id_offsets[r + 1] - id_offsets[r]) + 1 + len[r] + 1;
        if (len[r] < 0 || out_offsets[r + 1] - out_offsets[r] != need) {
            set_error("radian_fasta_records_host: record %d needs %lld bytes, its slot has %lld", r, (long long)need,
                      (long long)(out_offsets[r + 1] - out_offsets[r]));
            return RADIAN_E_ARG;
        }
    }
    unsigned nt = std::thread::hardware_concurrency();
    nt = nt < 1 ? 1 : nt > 16 ? 16 : nt;
    if (n_reads < 256) nt = 1;
    auto work = [&](int lo, int hi) {
        for (int r = lo; r < hi; ++r) {
            char *o = out + out_offsets[r];
            const int64_t idn = id_offsets[r + 1] - id_offsets[r];
            *o++ = '>';
            memcpy(o, ids + id_offsets[r], (size_t)idn);
            o += idn;
            *o++ = '\n';
            const uint8_t *s = seq + seq_offsets[r];
            for (int64_t i = len[r] - 1; i >= 0; --i) *o++ = bases[s[i] & 3];
            *o = '\n';
        }
    };
    if (nt == 1) {
        work(0, n_reads);
    } else {
        std::vector<std::thread> th;
        for (unsigned k = 0; k < nt; ++k) th.emplace_back(work, (int)((int64_t)n_reads * k / nt), (int)((int64_t)n_reads * (k + 1) / nt));
        for (auto &t : th) t.join();
    }
    return RADIAN_OK;
}
