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

// natural log of p x 2^kacc (a read's final score).  The value is split into mantissa and exponent
// first, so that the result does not depend on where the read happened to be rescaled.
static __device__ __noinline__ double final_log_score(double p, long long kacc)
{
    if (!(p > 0.0)) return -INFINITY;
    int e;
    const double m = frexp(p, &e);
    return log(m) + (double)(kacc + e) * 0.693147180559945309417;
}

// shared-memory loads by 32-bit shared address (the quiet loop of decode.cu keeps its addresses that way)
__device__ __forceinline__ double lds_f64(unsigned addr)
{
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ double lds_f64_volatile(unsigned addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ int lds_i32(unsigned addr)
{
    int v;
    asm("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ int4 lds_i4(unsigned addr)
{
    int4 v;
    asm("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

// max over the symbols c whose extension is a candidate of its own (bit 7 of byte c of km) of the high
// word of the table value r_c; row = the four float64 of a beam's extend-context
__device__ __forceinline__ int row_bound(const double *row, uint32_t km)
{
    const int4 ra = *reinterpret_cast<const int4 *>(row);      // r0 lo,hi r1 lo,hi
    const int4 rb = *reinterpret_cast<const int4 *>(row + 2);  // r2, r3
    return max(max(ra.y & (int)byte_sign_mask<0>(km), ra.w & (int)byte_sign_mask<1>(km)),
               max(rb.y & (int)byte_sign_mask<2>(km), rb.w & (int)byte_sign_mask<3>(km)));
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
// Record layout (doubles; LM: 18 per frame, no model: 10):
//   rec[0..4]  P0..P3, blank           rec[5]  entropy gate as 1.0 / 0.0
//   rec[6..9]  p/S (gate open) or P0..P3 again (closed)      rec[10] 1.0
//   rec[11]    S/2 (gate open; combine_dists' ((r + q)/2)*S is (r + q)*(S/2), which rounds
//              identically) or 1.0 (closed)
//   ints 24..27 (12..15 without the model): high words of P0..P3;  ints 28..31: of p/S;
//   ints 32..35 (int 10 without the model): words of the quiet-frame test, see record_ext.
// COMPACT (decode.cu) leaves the two arrays of high words out: 14 doubles per frame with the model
// (test words at ints 24..27), 6 without (test word at int 10).
// slack of the integer bound (see decode.cu): 2 * 0.0861 * 2^20 for p * P_c, one more 0.0861 for the
// max(r, q) * S bound of a gated extension; both minus the exponent bias of one float64 factor
constexpr int kSlackPlain = 181000 - 0x3ff00000;
constexpr int kSlackGated = 272000 - 0x3ff00000;

template <bool LM, bool EXT, bool COMPACT>
__device__ __forceinline__ void record_ext(double *rec, const double *v, const double *q, double S, bool gate)
{
    if (EXT) {
        int *ri = reinterpret_cast<int *>(rec);
        const int h0 = __double2hiint(v[0]), h1 = __double2hiint(v[1]), h2 = __double2hiint(v[2]), h3 = __double2hiint(v[3]);
        if (!COMPACT) *reinterpret_cast<int4 *>(ri + (LM ? 24 : 12)) = make_int4(h0, h1, h2, h3);
        // first-stage bound of decode.cu's quiet-frame test: the largest high word over all four
        // symbols, with the slack and the bias folded in
        const int hmaxP = max(max(h0, h1), max(h2, h3)) + kSlackPlain;
        if (LM) {
            // rec[6..9] and rec[11] are only meant for frames whose entropy gate is open; with the
            // gate closed they repeat P0..P3 and hold 1.0, and rec[10] is always 1.0: decode.cu's
            // copy emission (rcopy * rec[5] + rec[x]) * rec[y] then needs no branch (x, y per lane)
            rec[10] = 1.0;
            rec[11] = gate ? 0.5 * S : 1.0;
            if (!gate) {
#pragma unroll
                for (int i = 0; i < 4; ++i) rec[6 + i] = v[i];
            }
            const int g0 = __double2hiint(q[0]), g1 = __double2hiint(q[1]), g2 = __double2hiint(q[2]), g3 = __double2hiint(q[3]);
            if (!COMPACT) *reinterpret_cast<int4 *>(ri + 28) = make_int4(g0, g1, g2, g3);
            // {gate, hS, zP, zQ}: hS = high word of S + slack - 2 x bias; zP / zQ = first-stage words
            // of a plain / gated lane, which takes max(zQ, rmax) + hS.  With the gate closed zQ is
            // larger than any rmax and hS brings the sum back to zP (modulo 2^32).
            const int hS = __double2hiint(S) - 0x3ff00000 + kSlackGated;
            const int zQ = max(max(g0, g1), max(g2, g3));
            *reinterpret_cast<int4 *>(ri + (COMPACT ? 24 : 32)) = gate ? make_int4(1, hS, hmaxP, zQ)
                                                      : make_int4(0, (int)((unsigned)hmaxP - 0x7ff00000u), hmaxP, 0x7ff00000);
        } else {
            ri[10] = hmaxP;
        }
    }
}

// The entropy gate H_s > s_threshold in the reference's exact operation order (decode.py:73-76, 93), for
// the frames whose float32 estimate lands within 1e-4 of the threshold.  Out of line: four double
// logarithms are a few hundred instructions that the frame loop's instruction cache should never see.
static __device__ __noinline__ bool exact_gate(double q0, double q1, double q2, double q3, double s_thr)
{
    const double q[4] = {q0, q1, q2, q3};
    double H = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (q[i] > 0.0) H = __dadd_rn(H, __dmul_rn(q[i], log(q[i])));
    return -H > s_thr;
}
// float32 input: numpy >= 2 keeps the sum and the products in float32 and compares in float32
static __device__ __noinline__ bool exact_gate(float q0, float q1, float q2, float q3, double s_thr)
{
    const float q[4] = {q0, q1, q2, q3};
    float H = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (q[i] > 0.0f) H = __fadd_rn(H, __fmul_rn(q[i], __double2float_rn(log((double)q[i]))));
    return -H > (float)s_thr;
}

// SCALE (float64 input only): a row whose largest entry is below 2^-64 is multiplied by an exact
// power of two that brings it near 1; the exponent is returned and the caller adds it to the
// read's score exponent.  Every quantity the search derives from the row (S, p/S, the entropy
// gate, combine_dists) is invariant or scales exactly with it, so the result is the one the
// reference computes in the log domain, where magnitude is no issue (decode.py:16-17).  Rows of
// softmax outputs are never touched.  Returns k such that stored row = true row x 2^k.
template <bool LM, bool EXT = false, bool SCALE = false, bool COMPACT = false>
__device__ __forceinline__ int make_record(const double *raw, double s_thr, double *rec)
{
    double v[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) v[i] = raw[i];
    int k = 0;
    if (SCALE) {
        int emax = 0;
#pragma unroll
        for (int i = 0; i < 5; ++i) emax = max(emax, (__double2hiint(v[i]) >> 20) & 0x7ff);
        if (emax < 1023 - 64) {
            // subnormal rows first come up by 2^1000 (exact), the next rescale does the rest
            k = emax == 0 ? 1000 : 1022 - emax;
            const double sc = __hiloint2double((1023 + k) << 20, 0);
#pragma unroll
            for (int i = 0; i < 5; ++i) v[i] = __dmul_rn(v[i], sc);
        }
    }
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
        bool gate = Ha > (float)s_thr;
        if (!(fabsf(Ha - (float)s_thr) > 1e-4f)) gate = exact_gate(q[0], q[1], q[2], q[3], s_thr);
        rec[5] = gate ? 1.0 : 0.0;
        record_ext<LM, EXT, COMPACT>(rec, v, q, S, gate);
    } else {
        record_ext<LM, EXT, COMPACT>(rec, v, v, 0.0, false);
    }
    return k;
}

template <bool LM, bool EXT = false, bool SCALE = false, bool COMPACT = false>
__device__ __forceinline__ int make_record(const float *raw, double s_thr, double *rec)
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
        bool gate = Ha > (float)s_thr;
        if (!(fabsf(Ha - (float)s_thr) > 1e-4f)) gate = exact_gate(q[0], q[1], q[2], q[3], s_thr);
        rec[5] = gate ? 1.0 : 0.0;
        const double vd[4] = {(double)v[0], (double)v[1], (double)v[2], (double)v[3]};
        const double qd[4] = {(double)q[0], (double)q[1], (double)q[2], (double)q[3]};
        record_ext<LM, EXT, COMPACT>(rec, vd, qd, (double)S, gate);
    } else {
        const double vd[4] = {(double)v[0], (double)v[1], (double)v[2], (double)v[3]};
        record_ext<LM, EXT, COMPACT>(rec, vd, vd, 0.0, false);
    }
    return 0;  // float32-derived probabilities are >= 2^-149: no row scaling needed
}

}  // namespace radian
