// Device helpers shared by the decode kernels (decode.cu: beam widths <= 32, one beam per lane;
// decode_wide.cu: beam widths 33..128, several beams per lane).
#pragma once
#include <math.h>

#include "internal.h"

namespace radian {

constexpr unsigned kFull = 0xffffffffu;
constexpr uint16_t kPosInvalid = 0xffff;
constexpr int kNursery = 4096;  // arena nodes between two collections
constexpr unsigned kStallUnits = 512;  // ~0.54 s in units of 2^20 ns: a streamed transfer that does not
                                       // advance for this long is given up (see DecodeArgs::ready)

__device__ __forceinline__ unsigned timer_units()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return (unsigned)(t >> 20);
}

// all-ones if bit 7 of byte `B` of x is set, else zero (prmt with the sign-replicate selector bit)
template <int B>
__device__ __forceinline__ uint32_t byte_sign_mask(uint32_t x)
{
    uint32_t r;
    asm("prmt.b32 %0, %1, 0, %2;" : "=r"(r) : "r"(x), "n"(0x8888 + 0x1111 * B));
    return r;
}

__device__ __forceinline__ unsigned long long hash_step(unsigned long long h, int c)
{
    h = (h ^ (unsigned long long)(c + 1)) * 0x9E3779B97F4A7C15ull;
    return h ^ (h >> 29);
}

// Asynchronous global -> shared copies (LDGSTS): no register staging, completion awaited with
// cp_async_wait_all() by the issuing thread right before the data is needed.
template <int BYTES>
__device__ __forceinline__ void cp_async(void *smem_dst, const void *gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    if (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], %2;\n" ::"r"(d), "l"(gmem_src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// one posterior row (5 values of 4 or 8 bytes; rows are only element-aligned) -> smem
template <typename PT>
__device__ __forceinline__ void prefetch_row(PT *dst, const PT *post, long long frame)
{
    const PT *r = post + frame * 5;
#pragma unroll
    for (int i = 0; i < 5; ++i) cp_async<sizeof(PT)>(dst + i, r + i);
}

// Per-frame work shared by all beams of a read, done by the lane that loaded the frame.
// Reference: decode.py:135-138 (s_entropies), 67-76 (normalise, entropy), 54-55 (base sum and
// p/S of combine_dists).  float32 input follows the numpy>=2 promotion the pinned oracle uses.
// The entropy only feeds the comparison H > s_threshold (decode.py:93): a float32 estimate with
// the hardware log decides it unless it lands within 1e-4 of the threshold, in which case the
// reference's exact operation order is evaluated.
//
// EXT adds what decode.cu's quiet-frame test reads instead of the float64 values: the high words
// of P0..P3 (int4 at double index 12 with the model, 6 without), of q0..q3 (int4 at 14), and
// {gate, high word of S minus the exponent bias} (int2 at 16); rec[11] = S/2 (exact), so that
// combine_dists' ((r + q)/2)*S is one add and one multiply: (r + q)*(S/2) rounds identically.
template <bool LM, bool EXT>
__device__ __forceinline__ void record_ext(double *rec, const double *v, const double *q, double S, bool gate)
{
    if (EXT) {
        int *ri = reinterpret_cast<int *>(rec);
        *reinterpret_cast<int4 *>(ri + (LM ? 24 : 12)) =
            make_int4(__double2hiint(v[0]), __double2hiint(v[1]), __double2hiint(v[2]), __double2hiint(v[3]));
        if (LM) {
            rec[11] = 0.5 * S;
            *reinterpret_cast<int4 *>(ri + 28) =
                make_int4(__double2hiint(q[0]), __double2hiint(q[1]), __double2hiint(q[2]), __double2hiint(q[3]));
            *reinterpret_cast<int2 *>(ri + 32) = make_int2(gate ? 1 : 0, __double2hiint(S) - 0x3ff00000);
        }
    }
}

template <bool LM, bool EXT = false>
__device__ __forceinline__ void make_record(const double *raw, double s_thr, double *rec)
{
    double v[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) v[i] = raw[i];
#pragma unroll
    for (int i = 0; i < 5; ++i) rec[i] = v[i];
    if (LM) {
        const double S = __dadd_rn(__dadd_rn(__dadd_rn(v[0], v[1]), v[2]), v[3]);
        double q[4];
        float Ha = 0.0f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            q[i] = (S == 0.0) ? v[i] : v[i] / S;
            rec[6 + i] = q[i];
            const float qf = (float)q[i];
            if (qf > 0.0f) Ha -= qf * __logf(qf);
        }
        rec[10] = S;
        bool gate = Ha > (float)s_thr;
        if (!(fabsf(Ha - (float)s_thr) > 1e-4f)) {
            double H = 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (q[i] > 0.0) H = __dadd_rn(H, __dmul_rn(q[i], log(q[i])));
            gate = -H > s_thr;
        }
        rec[5] = gate ? 1.0 : 0.0;
        record_ext<LM, EXT>(rec, v, q, S, gate);
    } else {
        record_ext<LM, EXT>(rec, v, v, 0.0, false);
    }
}

template <bool LM, bool EXT = false>
__device__ __forceinline__ void make_record(const float *raw, double s_thr, double *rec)
{
    float v[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) v[i] = raw[i];
#pragma unroll
    for (int i = 0; i < 5; ++i) rec[i] = (double)v[i];
    if (LM) {
        const float S = __fadd_rn(__fadd_rn(__fadd_rn(v[0], v[1]), v[2]), v[3]);
        float q[4];
        float Ha = 0.0f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            q[i] = (S == 0.0f) ? v[i] : __fdiv_rn(v[i], S);
            rec[6 + i] = (double)q[i];
            if (q[i] > 0.0f) Ha -= q[i] * __logf(q[i]);
        }
        rec[10] = (double)S;
        bool gate = Ha > (float)s_thr;
        if (!(fabsf(Ha - (float)s_thr) > 1e-4f)) {
            float H = 0.0f;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (q[i] > 0.0f) H = __fadd_rn(H, __fmul_rn(q[i], __double2float_rn(log((double)q[i]))));
            gate = -H > (float)s_thr;
        }
        rec[5] = gate ? 1.0 : 0.0;
        const double vd[4] = {(double)v[0], (double)v[1], (double)v[2], (double)v[3]};
        const double qd[4] = {(double)q[0], (double)q[1], (double)q[2], (double)q[3]};
        record_ext<LM, EXT>(rec, vd, qd, (double)S, gate);
    } else {
        const double vd[4] = {(double)v[0], (double)v[1], (double)v[2], (double)v[3]};
        record_ext<LM, EXT>(rec, vd, vd, 0.0, false);
    }
}

}  // namespace radian
