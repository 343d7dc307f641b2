// Internal declarations shared by the translation units of libradian_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <vector>

#include "radian_b200.h"

namespace radian {

void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

#define RADIAN_CUDA(call)                                          \
    do {                                                           \
        cudaError_t e__ = (call);                                  \
        if (e__ != cudaSuccess) return ::radian::cuda_fail(e__, #call); \
    } while (0)

struct DeviceInfo {
    int sm_count;
    int max_smem_optin;
    int smem_per_sm;
};
int device_info(int device, DeviceInfo *out);
// this library's private stream-ordered memory pool on `device` (created on first use)
int keep_pool(int device, cudaMemPool_t *pool);
// One host entry point at a time per device: while a streamed decode kernel waits for its input,
// no other thread of this library may allocate, free or launch on that device.
std::mutex &host_mutex(int device);

}  // namespace radian

// One gate mask per r_threshold the table has been used with: bit i = entropy(row i) < threshold
// (decode.py:93).  Masks are immutable once built; `ready` is recorded behind the kernel that fills
// the mask and every consuming stream waits for it.
struct GateMask {
    double threshold;
    uint32_t *d_bits;
    cudaEvent_t ready;
    unsigned long long last_use;
};

struct radian_table {
    int L = 0;
    int device = 0;
    size_t rows = 0;
    double *d_rows = nullptr;     // rows x 4 float64 linear probabilities
    double *d_entropy = nullptr;  // rows float64, reference entropy() of each row (host libm)
    uint32_t *d_miss = nullptr;   // rows/32 words, bit = context absent from the model (KeyError when a
                                  // kept beam reaches it, decode.py:83); nullptr for a complete table
    int rcap = 0;                 // high word of the largest entry of d_rows
    mutable std::mutex mu;        // guards the mask cache below
    mutable std::vector<GateMask> gates;
    mutable unsigned long long tick = 0;
};

namespace radian {

// gate mask of `t` for r_threshold, built on `stream` if this threshold is new; `stream` is made to
// wait for the mask either way.  Calls with different thresholds on different streams do not
// disturb each other (a small LRU of masks).
int table_get_gate(const radian_table *t, double r_threshold, cudaStream_t stream, const uint32_t **d_bits);

struct DecodeArgs {
    const void *post;
    const int64_t *frame_offsets;
    const int32_t *order;
    int n_reads;
    int beam_width;
    int64_t max_frames;      // frames of the longest read (launch shape only)
    int64_t total_frames;    // frames of all reads, 0 = unknown (launch shape only)
    int excl_frames;         // > 0: reads of at least this many frames get a warp to themselves (set by decode_launch)
    const double *table;     // nullptr = LM off
    const uint32_t *gate;
    const uint32_t *miss;    // optional: contexts absent from the model (radian_table::d_miss)
    int L;
    int rcap;                // high word of the largest table entry: bound of a row that is still in flight
    double s_thr;
    uint8_t *out_seq;
    const int64_t *seq_offsets;
    int64_t *out_len;
    double *out_score;
    int32_t *out_status;
    unsigned long long *out_counters;
    uint32_t *arena;         // slots x (arena_cap + nursery) words: nodes, then forwarding scratch
    int arena_cap;
    int *queue;              // work-queue head, zeroed before launch; queue[1] = "transfer stalled, stop
                             // waiting" flag of streamed batches
    const int *ready;        // optional, 2 ints (8-byte aligned): reads (in queue order) whose posteriors
                             // have landed in HBM, as published by each of the two copy streams of the
                             // _host entry point while the kernel runs; a read is usable once both
                             // counters have passed it.  nullptr = everything is resident at launch.
};

struct DecodeLaunch {
    const void *kernel;
    size_t smem;             // dynamic shared memory per CTA
    int grid;
    int block;
    int groups_per_block;
};

int decode_pick(int device, int beam_width, bool lm, bool f64, bool count, bool stream, DecodeLaunch *out);
int decode_launch(const DecodeArgs &a, bool f64, int device, cudaStream_t stream);
int decode_max_slots(int device, int beam_width);
int64_t decode_arena_cap(int beam_width, int64_t max_frames, int64_t arena_nodes);
int decode_nursery();
// decode_wide.cu: kernel for beam widths 33..128; returns warps (= reads) per CTA
int wide_pick(int beam_width, bool lm, bool f64, bool count, const void **kernel, size_t *smem_bytes);

}  // namespace radian
