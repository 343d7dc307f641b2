// Device helpers shared by the decode kernels (decode.cu: beam widths <= 32, one beam per lane;
// decode_wide.cu: beam widths 33..128, several beams per lane).
#pragma once
#include <math.h>

#include "internal.h"

namespace radian {

constexpr unsigned kFull = 0xffffffffu;
constexpr uint16_t kPosInvalid = 0xffff;
constexpr int kNursery = 4096;  // arena nodes between two collections

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
template <bool LM>
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
    }
}

template <bool LM>
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
    }
}

}  // namespace radian
