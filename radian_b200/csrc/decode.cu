// CTC prefix beam search with gated RNA-model fusion, sm_100a.
//
// What it computes is radian/decode.py:100-212 of the reference (beam_search), per read:
//   frame loop 141-204, COPY 150-175, EXTEND 177-201, apply_rna_model 79-96,
//   combine_dists 52-64, frame entropy 135-138 (+67-76), final pick 207-210.
// How it computes it is new:
//   * one group of G lanes (8/16/32) owns a read; lane == beam, all per-beam state in registers;
//     a warp carries 32/G reads; a persistent grid pulls reads from an atomic queue.
//   * scores are kept in the LINEAR domain as float64, rescaled by an exact power of two whenever
//     the best beam falls below 2^-256 (exponent accumulated in an integer).  logaddexp becomes '+', '+ log p' becomes
//     '* p', and combine_dists is already linear, so the frame loop has no transcendental at
//     all; ordering by pr_total is unchanged because log is monotone.  One log per read at
//     the end gives the reference's log score.
//   * the T x 5 posterior rows are streamed through shared memory in tiles of G frames: each
//     lane loads one frame (prefetched one tile ahead), does the per-frame work that does not
//     depend on the beams (float64 conversion, base-sum, p/S, entropy gate) once, and the G
//     lanes then consume the records by broadcast reads.
//   * a labeling is identified by a 64-bit rolling hash + length (the reference's dict key is
//     the tuple itself); the only collision the algorithm can produce is copy(X) with
//     extend(parent(X), last(X)), found through a parent-lane pointer kept per beam; the child
//     computes the merged term itself from the parent's two scores and its own copy emission.
//   * most frames are QUIET: the rank order of the copies (kept as state) holds and no extension
//     can reach the worst copy, which an integer bound on the high words of the float64 values
//     proves without computing a single extension; such a frame is one vote and three score
//     updates per beam.  Only otherwise are the extension scores formed and, if needed, ranked.
//   * stable top-k: rank = #candidates with (score desc, dict insertion position asc) before
//     it, counted on the high words when they are all distinct (proved by the rank sum), else
//     exactly on the float64 bit patterns; candidates that cannot reach the beam (below the
//     worst copy) are pruned first.
//   * labelings live in a per-group back-pointer arena (parent<<2|symbol).  The live beams of a
//     read are combinations of its few most ambiguous positions, so their common ancestor stays
//     near the start of the read: nothing can be flushed early.  The arena is generational
//     instead: new nodes go to a small nursery; when it fills, the nodes still reachable from a
//     live beam are slid down onto the old generation (never revisited) and the rest is dropped.
//     Every lineage promotes each of its symbols once, so the old generation is bounded by
//     beam_width x decoded length.  The best labeling is read back by one walk at the end.
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>

#include "decode_common.cuh"

namespace radian {

#ifndef RADIAN_WARPS
#define RADIAN_WARPS 4
#endif
constexpr int kWarpsPerBlock = RADIAN_WARPS;
// resident CTAs per SM asked from ptxas.  A/B on B200 (profiles/r2_history.md): 5 CTAs = 20 warps at
// <= 96 registers beat 6 CTAs at <= 80 by 15-20 %: at 80 the frames that change the beam set spill.
#ifndef RADIAN_MIN_BLOCKS
#define RADIAN_MIN_BLOCKS 5
#endif

template <int G, bool LM, typename PT>
struct __align__(16) GroupSmem {
    // doubles per frame (decode_common.cuh, compact layout): P0..P4, gate, p/S, 1, S/2, then the
    // integer words of the quiet-frame test
    static constexpr int REC = LM ? 14 : 6;  // compact records (decode_common.cuh)
    double rec[G * REC];
    double row[LM ? G * 4 : 4];       // RNA table row of every lane's extend-context (cp.async target)
    PT raw[G * 5];                    // next tile of posterior rows, landed by cp.async
    double ex[G * 2];                 // {pr_total, pr_blank} of every lane before the frame (copy/extend merge)
    double ex2[G * 2];                // the same for the second frame of a pair (quiet loop, two frames per turn)
    double zero[2];                   // 0.0: what a beam without a live parent reads as its parent's score
                                      // (two of them: the arrays below are read with 16-byte loads)
    unsigned long long key[5 * G];    // candidate list: [0,G) copies by lane, [G,..) extensions
    uint32_t k32[2 * G];              // high words of the copies [0,G) and of the listed extensions [G,2G)
    uint32_t kill[G];                 // byte c of word l: extension (l,c) merged into a copy
    uint16_t pos[5 * G];              // dict insertion position of the candidate
    uint8_t src[5 * G];               // lane*4+c of an extension candidate
    uint8_t rnk[5 * G];               // rank of the candidate
    uint8_t newlist[G];               // candidate indices of the new beams, in list order
    // Per-beam state that only the frames that change the beam set touch (and the start and end of a
    // read): kept here instead of in registers, which the quiet loop then has for itself.
    unsigned long long c_h[G];        // hash of the labeling
    unsigned long long c_hp[G];       // hash of the parent labeling
    uint32_t c_ctx[G];                // last symbols, 2 bits each
    int c_len[G];                     // labeling length
    int c_node[G];                    // arena node of the last symbol
    int c_rank[G];                    // rank among the kept beams
};

// Streamed batches: how many reads (in queue order) have landed, or -1 when the transfer is given
// up.  Two copy streams publish alternating segments; both counters must have passed a read.
// Stalled-transfer guard: something on the host (another thread freeing memory, say) can hold the
// copies back until this very kernel has ended.  If the counter stands still for half a second
// everybody stops waiting; the reads not decoded keep their "not run" status and the host decodes
// them with a second launch once the data is there.  Kept out of line: it runs once per read at
// most, and inlined it costs the frame loop 8 % through register allocation.
__device__ __noinline__ int poll_arrivals(const int *ready, int *stop_flag, int idx, int *wait_seen, unsigned *wait_t0)
{
    const unsigned long long v = *(const volatile unsigned long long *)ready;
    const int r0 = (int)(unsigned)v, r1 = (int)(unsigned)(v >> 32);
    const int landed = r0 < r1 ? r0 : r1;
    if (landed > idx) return landed;
    const unsigned now = timer_units();
    if (*(const volatile int *)stop_flag != 0) return -1;
    if (landed != *wait_seen) {
        *wait_seen = landed;
        *wait_t0 = now;
    } else if (now - *wait_t0 > kStallUnits) {
        atomicExch(stop_flag, 1);
        return -1;
    }
    return landed;
}

// -DRADIAN_CHECKS: bounds checks of our own on every index the long way computes (arena, forwarding
// table, candidate lists); a violation prints where and traps, which fails the launch.  compute-sanitizer
// is closed on the build pool (profiles/r2_sanitizer_closed.txt); the test-suite is run once with this
// build instead (profiles/r2_checks_build.txt).
#ifdef RADIAN_CHECKS
#define RADIAN_ASSERT(c)                                                                       \
    do {                                                                                       \
        if (!(c)) {                                                                            \
            printf("radian check failed: %s (decode.cu:%d, block %d thread %d)\n", #c, __LINE__, \
                   (int)blockIdx.x, (int)threadIdx.x);                                         \
            __trap();                                                                          \
        }                                                                                      \
    } while (0)
#else
#define RADIAN_ASSERT(c)
#endif

// ---- rare, bulky pieces of the long way, out of line: the frame loop's instruction cache holds the
// usual route only (ncu on the bench launch: with everything inline the kernel was 80 KB of code and
// stalled 2.2 cycles per issued instruction on instruction fetch; profiles/r2_history.md)

// Nursery collection of the back-pointer arena (whole warps call it, once in thousands of frames).
// Returns the new top of the group's arena.
template <int G>
__device__ __noinline__ int collect_nursery(int *c_node, uint32_t *arena, uint32_t *fwd, int old_top, int top,
                                            bool run, bool alive, int li, int gshift)
{
    constexpr unsigned GBITS = (G == 32) ? kFull : ((1u << G) - 1u);
    const unsigned belowg = (1u << li) - 1u;
    // 1. mark nursery nodes reachable from a live beam (stop at the old generation or
    //    at a node somebody marked in an earlier step)
    const int node = c_node[li];
    int cur = node;
    bool walking = run && alive && cur >= old_top;
    while (__any_sync(kFull, walking)) {
        if (walking) {
            const uint32_t w = arena[cur];
            if (w >> 31) {
                walking = false;
            } else {
                arena[cur] = w | 0x80000000u;
                cur = (int)(w >> 2);
                walking = cur >= old_top;
            }
        }
        __syncwarp();
    }
    // 2. slide marked nodes down in index order (parents always precede children);
    //    fwd[] keeps the new index of every moved node for its children and the beams
    int iters = run ? (top - old_top + G - 1) / G : 0;
#pragma unroll
    for (int o = 16; o >= G; o >>= 1) {
        const int x = __shfl_xor_sync(kFull, iters, o);
        iters = x > iters ? x : iters;
    }
    int cnt = old_top;
    for (int k = 0; k < iters; ++k) {
        const int base = old_top + k * G;
        const int i = base + li;
        const uint32_t w = (run && i < top) ? arena[i] : 0u;
        const bool mk = (w >> 31) != 0;
        const unsigned bal = (__ballot_sync(kFull, mk) >> gshift) & GBITS;
        const int ni = cnt + __popc(bal & belowg);
        const int par = (int)((w & 0x7fffffffu) >> 2);
        int npar = par;
        if (mk && par >= old_top) {
            if (par >= base)
                npar = cnt + __popc(bal & ((1u << (par - base)) - 1u));
            else
                npar = (int)fwd[par - old_top];
        }
        __syncwarp();
        if (mk) {
            RADIAN_ASSERT(ni >= 0 && ni <= i && i - old_top < kNursery && npar < ni);
            arena[ni] = ((uint32_t)npar << 2) | (w & 3u);
            fwd[i - old_top] = (uint32_t)ni;
        }
        cnt += __popc(bal);
        __syncwarp();
    }
    if (run && alive && node >= old_top) c_node[li] = (int)fwd[node - old_top];
    __syncwarp();
    return cnt;
}

// exact rank of one candidate among the first `m` candidates of the group's list: (float64 bits desc,
// dict insertion position asc); bit 16 of the result: another candidate within 2^-40 of it
template <typename SM, bool COUNT>
__device__ __noinline__ int exact_rank(const SM &sm, int m, int self, unsigned long long k, int p)
{
    int cnt = 0;
    bool near = false;
    for (int j = 0; j < m; ++j) {
        const int pj = sm.pos[j];
        const unsigned long long kj = sm.key[j];
        cnt += (pj != kPosInvalid) && (kj > k || (kj == k && pj < p));
        // two candidates within 2^-40 of each other: a decision that the log-domain reference takes on
        // its own rounding noise
        if (COUNT) near = near || (j != self && pj != kPosInvalid && k != 0ull && kj - k + 4096ull < 8192ull);
    }
    return (cnt > 255 ? 255 : cnt) | (near ? 0x10000 : 0);
}

// all candidates of the list against each other, exactly (a beam that is still filling up has more
// extensions than lanes); ranks to sm.rnk, returns the near-tie flag
template <int G, typename SM, bool COUNT>
__device__ __noinline__ bool exact_rank_all(SM &sm, int m, int li)
{
    bool near = false;
    for (int idx = li; idx < m; idx += G) {
        const int p = sm.pos[idx];
        if (p != kPosInvalid) {
            const int r = exact_rank<SM, COUNT>(sm, m, idx, sm.key[idx], p);
            sm.rnk[idx] = (uint8_t)(r & 0xff);
            near = near || (r >> 16);
        }
    }
    return near;
}

// The COUNT instantiations also report, in the fourth counter of a read, (frames whose entropy gate
// H_s > s_threshold was open << 32) | frames that took the long way (see include/radian_b200.h)
#define RADIAN_STAT(x) if (COUNT) { x }

// (A/B on B200, profiles/r2_history.md: with one frame in eight taking the long way, a pair loop spends
// on discarded second frames what it saves on votes; it only pays for a warp that is alone on its
// scheduler.  Off by default.)
#ifndef RADIAN_PAIRS
#define RADIAN_PAIRS 0
#endif
// Two frames per turn: the second frame is formed from the first one's results before anybody
// knows whether the first was quiet, and one vote decides for both (the vote and the branch on it
// are the longest link of a frame's dependency chain).  If only the first of the pair was quiet it
// is committed alone and the loop ends at the second.
#define RADIAN_QUIET_PAIRS(IDLE)                                                                             \
    _Pragma("unroll 1") for (; it + 1 < nend; it += 2)                                                      \
    {                                                                                                        \
        const unsigned rb = q_rb0 + (unsigned)it * (unsigned)(REC * 8);                                      \
        constexpr unsigned RB2 = REC * 8;                                                                    \
        const double P4a = lds_f64(rb + 32), P4b = lds_f64(rb + RB2 + 32);                                   \
        double dla = lds_f64(rb + q_ox), dlb = lds_f64(rb + RB2 + q_ox);                                     \
        int za, zb;                                                                                          \
        bool fga = false, fgb = false;                                                                       \
        if (LM) {                                                                                            \
            const double ya = lds_f64(rb + q_oy), yb = lds_f64(rb + RB2 + q_oy);                             \
            const double ga = lds_f64(rb + 40), gb = lds_f64(rb + RB2 + 40);                                 \
            const int4 gia = lds_i4(rb + 96), gib = lds_i4(rb + RB2 + 96);                                   \
            dla = __dmul_rn(__dadd_rn(__dmul_rn(rcopy, ga), dla), ya);                                       \
            dlb = __dmul_rn(__dadd_rn(__dmul_rn(rcopy, gb), dlb), yb);                                       \
            za = gext ? max(gia.w, rmax) + gia.y : gia.z;                                                    \
            zb = gext ? max(gib.w, rmax) + gib.y : gib.z;                                                    \
            fga = gia.x != 0;                                                                                \
            fgb = gib.x != 0;                                                                                \
        } else {                                                                                             \
            za = lds_i32(rb + 40);                                                                           \
            zb = lds_i32(rb + RB2 + 40);                                                                     \
        }                                                                                                    \
        /* first frame */                                                                                    \
        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(ex_addr), "d"(ptot), "d"(pb) : "memory");      \
        double npnbA = __dmul_rn(pnb, dla);                                                                  \
        const double npbA = __dmul_rn(ptot, P4a);                                                            \
        double nptotA = __dadd_rn(npbA, npnbA);                                                              \
        __syncwarp();                                                                                        \
        {                                                                                                    \
            const double v = __dmul_rn(lds_f64_volatile(q_paddr), dla);                                      \
            npnbA = __dadd_rn(npnbA, v);                                                                     \
            nptotA = __dadd_rn(nptotA, v);                                                                   \
        }                                                                                                    \
        /* second frame, from the first one's results */                                                     \
        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(ex_addr + (unsigned)(G * 16)), "d"(nptotA), "d"(npbA) : "memory"); \
        double npnbB = __dmul_rn(npnbA, dlb);                                                                \
        const double npbB = __dmul_rn(nptotA, P4b);                                                          \
        double nptotB = __dadd_rn(npbB, npnbB);                                                              \
        __syncwarp();                                                                                        \
        {                                                                                                    \
            const double v = __dmul_rn(lds_f64_volatile(q_paddr2), dlb);                                     \
            npnbB = __dadd_rn(npnbB, v);                                                                     \
            nptotB = __dadd_rn(nptotB, v);                                                                   \
        }                                                                                                    \
        const uint32_t kA = (uint32_t)__double2hiint(nptotA), kB = (uint32_t)__double2hiint(nptotB);         \
        const uint32_t ksA = __shfl_sync(kFull, kA, q_succ), ksB = __shfl_sync(kFull, kB, q_succ);           \
        const uint32_t kwA = __shfl_sync(kFull, kA, last_lane), kwB = __shfl_sync(kFull, kB, last_lane);     \
        const bool quietA = (IDLE && q_idle) || (kA >= ksA + q_inc && kwA >= 0x00100000u &&                  \
                                                 __double2hiint(ptot) + za < (int)kwA);                      \
        const bool quietB = (IDLE && q_idle) || (kB >= ksB + q_inc && kwB >= 0x00100000u && (int)kA + zb < (int)kwB); \
        if (!__all_sync(kFull, quietA && quietB)) {                                                          \
            pair_broke = true;                                                                               \
            if (!__all_sync(kFull, quietA)) break;                                                           \
            if (COUNT && LM) {                                                                               \
                n_lookup += (unsigned)c_lookup;                                                              \
                if (fga) n_combine += (unsigned)c_combine;                                                   \
                if (fga) n_stage2 += 1ull << 32;                                                             \
            }                                                                                                \
            ptot = nptotA;                                                                                   \
            pnb = npnbA;                                                                                     \
            pb = npbA;                                                                                       \
            ++it;                                                                                            \
            break;                                                                                           \
        }                                                                                                    \
        if (COUNT && LM) {                                                                                   \
            n_lookup += 2u * (unsigned)c_lookup;                                                             \
            n_combine += (fga ? (unsigned)c_combine : 0u) + (fgb ? (unsigned)c_combine : 0u);                \
            n_stage2 += ((unsigned long long)(fga ? 1 : 0) + (fgb ? 1 : 0)) << 32;                           \
        }                                                                                                    \
        ptot = nptotB; /* (a dead lane's new values are zero as well) */                                     \
        pnb = npnbB;                                                                                         \
        pb = npbB;                                                                                           \
    }

// The quiet loop (see the comment where it is used).  IDLE: some group of the warp has no read to
// run (its lanes hold zeros and must not block the vote); a separate copy of the loop, so that the
// common one does not carry the flag.  Everything a lane needs besides its three scores comes from
// REFRESH(): q_rb0 + q_ox / q_oy = where this lane's copy emission is in a record, q_paddr = the
// score of the parent whose extension merges into this beam (or a zero), q_succ / q_inc = the order
// check.
#define RADIAN_QUIET_LOOP(IDLE)                                                                              \
    _Pragma("unroll 1") for (; it < nend; ++it)                                                             \
    {                                                                                                        \
        const unsigned rb = q_rb0 + (unsigned)it * (unsigned)(REC * 8);                                      \
        const double P4 = lds_f64(rb + 32);                                                                  \
        double dl_ = lds_f64(rb + q_ox);                                                                     \
        int z;                                                                                               \
        bool fgate_ = false;                                                                                 \
        if (LM) {                                                                                            \
            /* COPY emission (decode.py:150-175, 58-61): (r + p/S) * (S/2) with the gate open and a gated   \
               copy-context, P[last] otherwise; both as (rcopy * gate + x) * y, see decode_common.cuh */    \
            const double y_ = lds_f64(rb + q_oy);                                                            \
            const double g_ = lds_f64(rb + 40);                                                              \
            const int4 gi = lds_i4(rb + 96);                                                                 \
            dl_ = __dmul_rn(__dadd_rn(__dmul_rn(rcopy, g_), dl_), y_);                                       \
            z = gext ? max(gi.w, rmax) + gi.y : gi.z;                                                        \
            fgate_ = gi.x != 0;                                                                              \
        } else {                                                                                             \
            z = lds_i32(rb + 40);                                                                            \
        }                                                                                                    \
        double npnb = __dmul_rn(pnb, dl_); /* the empty labeling and dead lanes have pnb == 0 */             \
        const double npb = __dmul_rn(ptot, P4);                                                              \
        double nptot = __dadd_rn(npb, npnb);                                                                 \
        /* MERGE with the parent's extension by my last symbol (see the slow frame) */                       \
        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(ex_addr), "d"(ptot), "d"(pb) : "memory");      \
        __syncwarp();                                                                                        \
        {                                                                                                    \
            const double v = __dmul_rn(lds_f64_volatile(q_paddr), dl_);                                      \
            npnb = __dadd_rn(npnb, v);                                                                       \
            nptot = __dadd_rn(nptot, v);                                                                     \
        }                                                                                                    \
        const uint32_t kc32 = (uint32_t)__double2hiint(nptot);                                               \
        const uint32_t ksucc = __shfl_sync(kFull, kc32, q_succ);                                             \
        const uint32_t kworst = __shfl_sync(kFull, kc32, last_lane);                                         \
        const bool quiet = (IDLE && q_idle) || (kc32 >= ksucc + q_inc && kworst >= 0x00100000u &&            \
                                                __double2hiint(ptot) + z < (int)kworst);                     \
        if (!__all_sync(kFull, quiet)) break;                                                                \
        if (COUNT && LM) {                                                                                   \
            n_lookup += (unsigned)c_lookup;                                                                  \
            if (fgate_) n_combine += (unsigned)c_combine;                                                    \
            if (fgate_) n_stage2 += 1ull << 32;                                                              \
        }                                                                                                    \
        ptot = nptot; /* (a dead lane's new values are zero as well) */                                      \
        pnb = npnb;                                                                                          \
        pb = npb;                                                                                            \
    }

// Control flow is warp-uniform everywhere: a warp carries 32/G reads, and every branch that
// contains a warp collective is taken by all of them together (decided by a full-mask vote), so
// all shuffles and votes use the full mask and each group extracts its own lanes' bits.  Sub-warp
// masks would make the compiler emit a convergence check per collective and serialise the groups.
// STREAM: the batch arrives while the kernel runs (DecodeArgs::ready); a separate instantiation,
// because the mere presence of the arrival polling costs the frame loop of resident launches 8 %
// (register allocation).
template <int G, bool LM, typename PT, bool COUNT, bool STREAM>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, RADIAN_MIN_BLOCKS)
decode_kernel(const DecodeArgs a)
{
    constexpr int GPW = 32 / G;  // groups (reads) per warp
    constexpr int REC = GroupSmem<G, LM, PT>::REC;
    constexpr unsigned GBITS = (G == 32) ? kFull : ((1u << G) - 1u);
    constexpr bool F64 = sizeof(PT) == 8;
    __shared__ GroupSmem<G, LM, PT> smem[kWarpsPerBlock * GPW];
    // streamed batches: last value of the arrival counter a waiting group saw, and since when
    // (stalled-transfer guard); kept out of GroupSmem, whose size the frame loop's addressing likes
    __shared__ int wait_seen[kWarpsPerBlock * GPW];
    __shared__ unsigned wait_t0[kWarpsPerBlock * GPW];

    const int lane = threadIdx.x & 31;
    const int li = lane % G;
    const int gw = lane / G;
    const int gshift = gw * G;
    const unsigned belowg = (1u << li) - 1u;  // lanes of my group below me, group-relative bits
    const int gib = (threadIdx.x >> 5) * GPW + gw;  // group in block
    GroupSmem<G, LM, PT> &sm = smem[gib];
    // shared-memory address of this lane's extension scores, pinned in a register (the compiler
    // would otherwise rebuild it from the lane and group indices every frame)
    if (STREAM && li == 0) wait_seen[gib] = -2;
    unsigned ex_addr = (unsigned)__cvta_generic_to_shared(&sm.ex[li * 2]);
    asm volatile("" : "+r"(ex_addr));
    const int slot = blockIdx.x * (kWarpsPerBlock * GPW) + gib;
    // votes: bits of my group's lanes, group-relative
#define GBALLOT(p) ((__ballot_sync(kFull, (p)) >> gshift) & GBITS)

    const int bw = a.beam_width;
    const int L = a.L;
    const uint32_t ctx_mask = LM ? (uint32_t)((1ull << (2 * L)) - 1ull) : 0u;
    const int cap = a.arena_cap;
    uint32_t *const arena = a.arena + (size_t)slot * (size_t)(cap + kNursery);
    uint32_t *const fwd = arena + cap;

    // ---- per-beam (lane) state; a dead lane keeps all three probabilities at zero
    double ptot = 0.0, pnb = 0.0, pb = 0.0;
    int plane = -1, last = 0;  // (hash, context, length, arena node and rank live in sm.c_*)
    int prep = 0;        // 1 if the live parent (plane) ends in the same symbol as this beam
    int rmax = 0;  // max high word of the unmerged entries of this beam's table row (row_bound), or a.rcap
    bool alive = false;
    double rcopy = 0;  // table value of this beam's last symbol in its copy-context; the row of the
                       // extend-context lives in sm.row[li*4..]
    bool gext = false, gcopy = false;
    bool rmax_prov = false;  // rmax is the table-wide bound: this beam's row was in flight when it was set
    int succ = 0;        // absolute lane of the beam ranked right after this one (own lane: none)
    bool tie_ok = false; // a successor with exactly my score is still in the right place (its insertion position is later)
    // Two copies next to each other in the order may agree in their high words for thousands of frames (two
    // readings of an old ambiguous position, multiplied by the same factors ever since).  The quiet loop only
    // compares high words; for such a pair, once the long way has confirmed the order on all 64 bits, it
    // accepts equal high words (hitie).  The low words could cross unseen while the high words stay equal; the
    // next frame that takes the long way re-checks all 64 bits, nothing in between uses the order, and the end
    // of the read picks its best beam exactly.
    bool hitie = false;
    uint32_t km = 0;     // byte c = 0x80: this lane holds a beam and its extension by c is a candidate
                         // of its own (not merged into a live child's copy); 0 for a dead lane
    // ---- per-read (group-uniform) state
    int read = -1, top = 0, old_top = 0, na = 0, status = 0, first_lane = 0, last_lane = 0;
    int T = 0, t = 0, pend = -1;  // pend: queue ticket of a read that has not landed yet
    long long kacc = 0;
    const PT *rp = (const PT *)a.post;
    unsigned long long n_lookup = 0, n_combine = 0, n_tie = 0, n_stage2 = 0;
#ifdef RADIAN_READ_TIMES
    unsigned long long t_read0 = 0;
#endif
    bool active = true;
    // ---- what the quiet loop reads besides the scores; derived from the state above by REFRESH()
    // whenever the beam set, the order or the read changes
    bool q_idle = true;
    const unsigned q_rb0 = (unsigned)__cvta_generic_to_shared(&sm.rec[0]);
    const unsigned q_zero = (unsigned)__cvta_generic_to_shared(&sm.zero[0]);
    unsigned q_paddr = q_zero, q_paddr2 = q_zero, q_ox = 0, q_oy = 80, q_inc = 0;
    int q_succ = lane;
    int c_lookup = 0, c_combine = 0;  // COUNT: lm[context] reads / combine_dists calls of a gated frame
    if (li == 0) sm.zero[0] = sm.zero[1] = 0.0;
#define REFRESH()                                                                                          \
    do {                                                                                                   \
        const bool run_ = active && read >= 0 && status == 0;                                              \
        const bool full_ = na >= bw;                                                                       \
        q_idle = !run_;                                                                                    \
        q_paddr = (run_ && alive && plane >= 0) ? (unsigned)__cvta_generic_to_shared(&sm.ex[plane * 2 + prep]) : q_zero; \
        q_paddr2 = (run_ && alive && plane >= 0) ? (unsigned)__cvta_generic_to_shared(&sm.ex2[plane * 2 + prep]) : q_zero; \
        /* copy emission: rec[6 + last] and rec[11] for a gated copy-context, rec[last] and rec[10] otherwise */ \
        q_ox = (unsigned)(((LM && gcopy) ? 6 + last : last) * 8);                                          \
        q_oy = (LM && gcopy) ? 88u : 80u;                                                                  \
        /* order check kc32 >= k(succ) + inc: strict for a beam with a successor, void for the last one    \
           (succ == lane) and not strict for a pair that shares its high word (hitie, below); a beam with  \
           room left is never quiet: every extension is a candidate */                                     \
        q_succ = full_ ? succ : lane;                                                                      \
        q_inc = (full_ && (succ == lane || hitie)) ? 0u : 1u;                                              \
        if (COUNT && LM) {                                                                                 \
            const int len_ = sm.c_len[li];                                                                \
            const bool lc_ = run_ && alive && len_ >= L + 1, le_ = run_ && alive && len_ >= L;               \
            c_lookup = __popc(GBALLOT(lc_)) + __popc(GBALLOT(le_));                                        \
            c_combine = __popc(GBALLOT(lc_ && gcopy)) + __popc(GBALLOT(le_ && gext));                      \
        }                                                                                                  \
    } while (0)
    // the read cannot be finished: its status is reported, its remaining frames are skipped and it
    // leaves no beam behind (an idle group must look "in order, nothing competing")
#define GIVE_UP(code)           \
    do {                        \
        status = (code);        \
        run = false;            \
        alive = false;          \
        km = 0u;                \
        succ = lane;            \
        ptot = pnb = pb = 0.0;  \
    } while (0)
    // best beam outside [2^300, 2^900): everything times an exact power of two, back to 2^600
#define RESCALE_CHECK()                                                                                    \
    do {                                                                                                   \
        const int exb_ = (__shfl_sync(kFull, __double2hiint(ptot), first_lane) >> 20) & 0x7ff;             \
        if (run && (unsigned)(exb_ - (1023 + 300)) >= 600u) {                                              \
            const int k1_ = exb_ == 0 ? 1000 : 1023 - exb_; /* a subnormal best first comes up by 2^1000 */ \
            const double s1_ = __hiloint2double((1023 + k1_) << 20, 0);                                    \
            const double s2_ = __hiloint2double((1023 + 600) << 20, 0);                                    \
            ptot = __dmul_rn(__dmul_rn(ptot, s1_), s2_);                                                   \
            pnb = __dmul_rn(__dmul_rn(pnb, s1_), s2_);                                                     \
            pb = __dmul_rn(__dmul_rn(pb, s1_), s2_);                                                       \
            kacc -= k1_ + 600;                                                                             \
        }                                                                                                  \
    } while (0)

    while (true) {
        // ------------------------------------------------------------ fetch a read
        {
            // EXCLUSIVE reads.  A warp carries 32/G reads and every one of them pays for the frames in
            // which one of the others changes its beam set; a read that is alone in its warp runs
            // about twice as fast.  For a batch that is finished when its longest reads are (the
            // host decides: a.excl_frames > 0), reads of at least that many frames -- they are at the
            // head of the longest-first queue -- are therefore only started by the first group of an
            // empty warp, and nothing else is started in that warp until they are done.
            bool may = true;
            if (!STREAM && a.excl_frames > 0 && GPW > 1) {
                const bool excl_now = __any_sync(kFull, read >= 0 && T >= a.excl_frames);
                const bool busy = __any_sync(kFull, read >= 0);
                int head_T = 0;
                if (lane == 0) {
                    const int qh = *(volatile const int *)a.queue;
                    if (qh < a.n_reads) {
                        const int r = a.order ? a.order[qh] : qh;
                        head_T = (int)(a.frame_offsets[r + 1] - a.frame_offsets[r]);
                    }
                }
                head_T = __shfl_sync(kFull, head_T, 0);
                if (excl_now)
                    may = false;
                else if (head_T >= a.excl_frames)
                    may = gw == 0 && !busy;
            }
            const bool want = active && read < 0 && may;
            int idx = pend;
            if (want && li == 0 && idx < 0) idx = atomicAdd(a.queue, 1);
            idx = __shfl_sync(kFull, idx, gshift);
            // streamed batch: the posteriors arrive over PCIe in queue order while the kernel runs
            // (a.ready = reads published by the copy stream so far).  A group whose read is still
            // in flight keeps its queue ticket and polls again later; it must not block the other
            // reads of its warp.
            int landed = 0x7fffffff;
#ifndef RADIAN_RESIDENT_KEEPS_POLL
#define RADIAN_RESIDENT_KEEPS_POLL 1
#endif
            if (STREAM) {
                if (want && li == 0 && idx < a.n_reads) landed = poll_arrivals(a.ready, a.queue + 1, idx, &wait_seen[gib], &wait_t0[gib]);
                landed = __shfl_sync(kFull, landed, gshift);
            } else if (RADIAN_RESIDENT_KEEPS_POLL && a.ready != nullptr) {
                // never taken (the host picks the STREAM instantiation whenever `ready` is set);
                // kept because the frame loop of the resident kernel compiles 3 % faster with it
                if (want && li == 0 && idx < a.n_reads) {
                    const unsigned long long v = *(const volatile unsigned long long *)a.ready;
                    const int r0 = (int)(unsigned)v, r1 = (int)(unsigned)(v >> 32);
                    landed = r0 < r1 ? r0 : r1;
                }
                landed = __shfl_sync(kFull, landed, gshift);
            }
            if (want) {
                if (idx >= a.n_reads || (STREAM && landed < 0)) {
                    active = false;
                    pend = -1;
                } else if ((STREAM || RADIAN_RESIDENT_KEEPS_POLL) && landed <= idx) {
                    pend = idx;
                } else {
                    pend = -1;
                    if (STREAM || (RADIAN_RESIDENT_KEEPS_POLL && a.ready != nullptr)) __threadfence();
                    read = a.order ? a.order[idx] : idx;
                    RADIAN_ASSERT(read >= 0 && read < a.n_reads);
                    const long long foff = a.frame_offsets[read];
                    T = (int)(a.frame_offsets[read + 1] - foff);
                    rp = (const PT *)a.post + foff * 5;
                    t = 0;
                    // initial beam: the empty labeling, pr_blank = pr_total = log 1 (decode.py:128-132)
                    alive = (li == 0);
                    ptot = alive ? 1.0 : 0.0;
                    pb = ptot;
                    pnb = 0.0;
                    sm.c_h[li] = 0x243F6A8885A308D3ull;
                    sm.c_hp[li] = 0;
                    sm.c_ctx[li] = 0;
                    sm.c_len[li] = 0;
                    sm.c_node[li] = 0;
                    sm.c_rank[li] = 0;
                    plane = -1;
                    prep = 0;
                    rmax = 0;
                    rmax_prov = false;
                    last = 0;
                    gext = gcopy = false;
                    succ = lane;
                    tie_ok = false;
                    hitie = false;
                    km = alive ? 0x80808080u : 0u;
                    first_lane = gshift;
                    last_lane = gshift;
                    top = 1;  // node 0 = the empty labeling
                    old_top = 1;
                    na = 1;
                    status = 0;
                    kacc = 0;
                    n_lookup = n_combine = n_tie = n_stage2 = 0;
#ifdef RADIAN_READ_TIMES
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_read0));
#endif
                }
            }
        }
        if (!__any_sync(kFull, active)) break;
        const bool live = active && read >= 0;  // this group has a read to run
        if (!__any_sync(kFull, live)) {
            __nanosleep(1000);  // everything this warp could run is still in flight
            continue;
        }

        // frames every live group of this warp can run before one of them finishes its read
        int rem = live ? (T - t) : 0x7fffffff;
        int nrun = rem;
#pragma unroll
        for (int g = 0; g < GPW; ++g) {
            const int x = __shfl_sync(kFull, rem, g * G);
            nrun = x < nrun ? x : nrun;
        }
        // a group waiting for its read gets another look at the flag after a bounded stretch
        if ((STREAM || RADIAN_RESIDENT_KEEPS_POLL) && __any_sync(kFull, active && !live) && nrun > 256) nrun = 256;
        // (re)prime the frame tiles so that all groups of the warp refill at the same iterations
        const int tb = t;
        __syncwarp();
        if (live && tb + li < T) prefetch_row(&sm.raw[li * 5], rp, tb + li);

        // (values the quiet loop reads per lane; they change only when the beam set does: REFRESH)
        REFRESH();

        for (int it0 = 0; it0 < nrun; it0 += G) {
            bool run = live && status == 0;  // group-uniform
            const int nend = (nrun - it0) < G ? (nrun - it0) : G;  // frames of this tile the warp runs
            // -------------------------------------------------------- tile refill
            {
                cp_async_wait_all();
                __syncwarp();
                int kf = 0;
                if (run && tb + it0 + li < T) kf = make_record<LM, true, F64, true>(&sm.raw[li * 5], a.s_thr, &sm.rec[li * REC]);
                __syncwarp();
                if (run && tb + it0 + G + li < T) prefetch_row(&sm.raw[li * 5], rp, tb + it0 + G + li);
                if (F64) {
                    // exponents taken out of tiny rows (make_record): added up for the frames that are
                    // going to be consumed; stored score = true score x 2^-kacc
                    kf = li < nend ? kf : 0;
                    if (__any_sync(kFull, kf != 0)) {
#pragma unroll
                        for (int o = G / 2; o > 0; o >>= 1) kf += __shfl_xor_sync(kFull, kf, o);
                        kacc -= kf;
                    }
                }
            }

            // -------------------------------------------------------- nursery collection
            // checked once per tile: a frame adds at most G nodes per read, a tile at most G*G
            if (__any_sync(kFull, run && (top + G * G > old_top + kNursery || top + G * G > cap))) {
                // every running group of the warp collects (early collection is harmless)
                const int cnt = collect_nursery<G>(sm.c_node, arena, fwd, old_top, top, run, alive, li, gshift);
                if (run) {
                    old_top = cnt;
                    top = cnt;
                    if (top + G * G > cap) GIVE_UP(RADIAN_READ_TRIE_OVERFLOW);
                }
                REFRESH();
            }

            // RESCALE by an exact power of two when the best beam (the largest value of the group)
            // has left [2^300, 2^900): back to 2^600.  Checked once per tile here and again before
            // every frame that is not quiet; in between, the quiet test itself refuses a frame in
            // which a kept beam is not a normal number, so nothing is ever lost to underflow.
            RESCALE_CHECK();

            int it = 0;
            while (true) {
                // ==================================================== QUIET frames
                // The common case: in every group of the warp the order of the copies holds strictly,
                // the beam is full and no extension, merged ones excepted, can reach the worst copy.
                // The extensions are not computed for this: with h(x) = high word of the float64 x,
                // (h(x) >> 20) - 1023 + mantissa fraction is a lower bound of log2 x that is short by at
                // most 0.0861, so
                //   h(p) + h(d) - bias + slack < h(worst)  implies  p*d < worst
                // for slack >= 2 * 0.0861 * 2^20.  The emission d of an extension is P_c, or with the
                // model ((r_c + q_c)/2) * S <= max(r_c, q_c) * S (one more 0.0861).  P_c and q_c are
                // bounded by the largest of the four symbols (one word per frame, prepared with the
                // record; measured on the bench workload, taking the largest over the unmerged
                // symbols only would keep 0.1 % more frames quiet); the table part of the bound, max
                // over the unmerged symbols of h(r_c), is a per-beam constant (rmax).
                // Such a frame is one vote and three score updates per beam; the loop carries
                // nothing but the three scores.
                bool pair_broke = false;
                if (__any_sync(kFull, q_idle)) {
                    RADIAN_QUIET_LOOP(true)
                } else if (RADIAN_PAIRS) {
                    RADIAN_QUIET_PAIRS(false)
                    // (a pair loop that stopped at a frame which is not quiet leaves it to the long way;
                    // one that ran out of pairs leaves at most the last frame of the tile)
                    if (it + 1 == nend && !pair_broke) {
                        RADIAN_QUIET_LOOP(false)
                    }
                } else {
                    RADIAN_QUIET_LOOP(false)
                }
                if (it >= nend) break;

                // ==================================================== one frame the long way
                // Everything is formed again from the scores before the frame (the quiet loop
                // committed nothing for it).
                const bool av = alive && run;  // this lane holds a beam that takes part in this frame
                int rank = sm.c_rank[li];
                RADIAN_STAT(n_stage2 += 1;)
                if (LM && __any_sync(kFull, av && rmax_prov)) {
                    // table bounds that were provisional (the row had not landed when the beam was
                    // created): take the real ones now
                    cp_async_wait_all();
                    if (av && rmax_prov) {
                        rmax = row_bound(&sm.row[li * 4], km);
                        rmax_prov = false;
                    }
                }
                RESCALE_CHECK();
                const double *rec = &sm.rec[it * REC];
                const int *reci = reinterpret_cast<const int *>(rec);
                const double P4 = rec[4];
                bool fgate = false;
                if (LM) fgate = reci[24] != 0;
                // (a dead lane computes on stale flags; all its scores are zero and stay zero)
                if (COUNT && LM) {
                    if (fgate && run) n_stage2 += 1ull << 32;
                    const int len_c = sm.c_len[li];
                    const bool lm_copy = av && len_c >= L + 1;  // decode.py:157
                    const bool lm_ext = av && len_c >= L;       // decode.py:180
                    n_lookup += __popc(GBALLOT(lm_copy)) + __popc(GBALLOT(lm_ext));
                    n_combine += __popc(GBALLOT(lm_copy && gcopy && fgate)) + __popc(GBALLOT(lm_ext && gext && fgate));
                }

                // COPY (decode.py:150-175).  (Everything is formed again here; handing the quiet loop's
                // values over instead was 7 % slower on B200: eight more registers live across the loop.)
                // the empty labeling and dead lanes have pnb == 0, so their copy needs no special case
                double dl_ = rec[last];
                // (gcopy implies len >= L+1 and gext implies len >= L: both are set when the beam is created)
                if (LM && gcopy && fgate) dl_ = __dmul_rn(__dadd_rn(rcopy, rec[6 + last]), rec[11]);  // decode.py:58-61
                double npnb = __dmul_rn(pnb, dl_);
                const double npb = __dmul_rn(ptot, P4);
                double nptot = __dadd_rn(npb, npnb);
                // a kept beam whose new score is not a normal number although its factors are not zero:
                // the spread of the beam exceeds what the rescaled float64 scores can express
                bool lost = av && ((ptot != 0.0 && P4 != 0.0) || (pnb != 0.0 && dl_ != 0.0));

                // MERGE copy(X) with extend(parent(X), last(X)): same dict key in the reference.  The
                // parent's extension by last(X) is (pr_blank or pr_total of the parent) x the emission
                // of last(X) in the parent's extend-context, and that emission is this beam's own
                // copy emission dl_ (same context, same gate, same table value), so only the parent's
                // two scores travel through shared memory.  Which pairs merge only changes when the
                // beam set changes: the pairing (plane, prep, km) is state.
                asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(ex_addr), "d"(ptot), "d"(pb) : "memory");
                __syncwarp();
                if (av && plane >= 0) {
                    const double pv = sm.ex[plane * 2 + prep];
                    const double v = __dmul_rn(pv, dl_);
                    npnb = __dadd_rn(npnb, v);
                    nptot = __dadd_rn(nptot, v);
                    lost = lost || (pv != 0.0 && dl_ != 0.0);
                }
                lost = lost && (uint32_t)__double2hiint(nptot) < 0x00100000u;

                // SELECT the best beam_width candidates (decode.py:145, 35-39).
                // The rank order is kept as state (succ = lane of the next-ranked beam) and re-validated
                // with one compare per beam; an extension matters only if it is not below the worst
                // copy of a full beam.
                const unsigned long long kcopy = (unsigned long long)__double_as_longlong(nptot);
                const uint32_t kc32 = (uint32_t)(kcopy >> 32);  // zero on a dead lane: its scores are zero
                const uint32_t kworst = __shfl_sync(kFull, kc32, last_lane);
                const bool prune = (na >= bw);
                bool ranks_changed = false;

                // EXTEND (decode.py:177-201)
                const double2 P01 = *reinterpret_cast<const double2 *>(rec);
                const double2 P23 = *reinterpret_cast<const double2 *>(rec + 2);
                double d0 = P01.x, d1 = P01.y, d2 = P23.x, d3 = P23.y;
                if (LM && gext && fgate) {
                    const double2 q01 = *reinterpret_cast<const double2 *>(rec + 6);
                    const double2 q23 = *reinterpret_cast<const double2 *>(rec + 8);
                    const double Sh = rec[11];
                    const double2 r01 = *reinterpret_cast<const double2 *>(&sm.row[li * 4]);
                    const double2 r23 = *reinterpret_cast<const double2 *>(&sm.row[li * 4 + 2]);
                    d0 = __dmul_rn(__dadd_rn(r01.x, q01.x), Sh);
                    d1 = __dmul_rn(__dadd_rn(r01.y, q01.y), Sh);
                    d2 = __dmul_rn(__dadd_rn(r23.x, q23.x), Sh);
                    d3 = __dmul_rn(__dadd_rn(r23.y, q23.y), Sh);
                }
                // a repeated symbol continues only paths that ended in a blank (decode.py:192-195); for
                // the empty labeling (last = 0 by convention) pb == ptot, so the rule is harmless there
                const double e0 = __dmul_rn(last == 0 ? pb : ptot, d0);
                const double e1 = __dmul_rn(last == 1 ? pb : ptot, d1);
                const double e2 = __dmul_rn(last == 2 ? pb : ptot, d2);
                const double e3 = __dmul_rn(last == 3 ? pb : ptot, d3);
                const unsigned long long ks64 = __shfl_sync(kFull, kcopy, succ);
                const bool order_ok = GBALLOT(kcopy > ks64 || (kcopy == ks64 && tie_ok) || succ == lane) == GBITS;
                // worst copy of the group: the last lane of the order when the order still holds.
                // With room left in the beam every extension is a candidate (threshold 1: keys are
                // or-ed with 1 so that a zero-probability extension of a live beam still counts).
                uint32_t tau = (prune && kworst > 1u) ? kworst : 1u;
                // extension keys, zeroed where the extension is merged into a child or the lane is dead
                const uint32_t x0 = ((uint32_t)__double2hiint(e0) | 1u) & byte_sign_mask<0>(km);
                const uint32_t x1 = ((uint32_t)__double2hiint(e1) | 1u) & byte_sign_mask<1>(km);
                const uint32_t x2 = ((uint32_t)__double2hiint(e2) | 1u) & byte_sign_mask<2>(km);
                const uint32_t x3 = ((uint32_t)__double2hiint(e3) | 1u) & byte_sign_mask<3>(km);
                const uint32_t xmax = max(max(x0, x1), max(x2, x3));
                if (!prune) {
                    // with room in the beam an extension that underflowed would be ranked as a zero
                    const double pe = ptot != 0.0 ? 1.0 : 0.0;  // (pb != 0 implies ptot != 0)
                    lost = lost || (av && pe != 0.0 &&
                                    ((x0 > 1u && x0 < 0x00100000u && d0 != 0.0 && (last == 0 ? pb : ptot) != 0.0) ||
                                     (x1 > 1u && x1 < 0x00100000u && d1 != 0.0 && (last == 1 ? pb : ptot) != 0.0) ||
                                     (x2 > 1u && x2 < 0x00100000u && d2 != 0.0 && (last == 2 ? pb : ptot) != 0.0) ||
                                     (x3 > 1u && x3 < 0x00100000u && d3 != 0.0 && (last == 3 ? pb : ptot) != 0.0)));
                }
                if (__any_sync(kFull, lost)) {
                    if (GBALLOT(lost) != 0u && run) GIVE_UP(RADIAN_READ_RANGE);
                    REFRESH();
                    continue;  // the other groups of the warp take this frame again: nothing was committed
                }
                // ---- which groups have something to decide
                bool tied = false;  // two copies share a high word: their order needs all 64 bits
                int base = rank;    // rank of my copy among the copies (unchanged while the order holds)
                if (__any_sync(kFull, !order_ok)) {
                    // some group's copies changed order: its worst copy is the minimum over the lanes,
                    // and the copies are ranked among themselves on the high words
                    uint32_t tmin = av ? kc32 : 0xffffffffu;
#pragma unroll
                    for (int o = G / 2; o > 0; o >>= 1) {
                        const uint32_t x = __shfl_xor_sync(kFull, tmin, o);
                        tmin = x < tmin ? x : tmin;
                    }
                    if (!order_ok) tau = (prune && tmin > 1u) ? tmin : 1u;
                    sm.k32[li] = kc32;
                    __syncwarp();
                    int cc = 0;
                    const uint4 *kv = reinterpret_cast<const uint4 *>(sm.k32);
#pragma unroll
                    for (int j = 0; j < G / 4; ++j) {
                        const uint4 k4 = kv[j];
                        cc += (k4.x > kc32) + (k4.y > kc32) + (k4.z > kc32) + (k4.w > kc32);
                    }
                    // all high words distinct <=> the ranks add up (any tie makes the sum fall short)
                    int ssum = av ? cc : 0;
#pragma unroll
                    for (int o = G / 2; o > 0; o >>= 1) ssum += __shfl_xor_sync(kFull, ssum, o);
                    if (!order_ok) {
                        base = cc;
                        tied = ssum != na * (na - 1) / 2;
                    }
                }
                const bool full = GBALLOT(run && xmax >= tau) != 0u;  // my group has extensions that compete

                if (!__any_sync(kFull, full || !order_ok)) {
                    // nothing competes after all (the integer bound is conservative) and the order
                    // holds: (a dead lane's new values are zero as well)
                    ptot = nptot;
                    pnb = npnb;
                    pb = npb;
                } else {
                    // Dict insertion position of my copy (decode.py:148-201: parents in rank order, copy
                    // before its own extensions; a merged entry keeps the earlier of its two positions).
                    // It decides between candidates whose scores are equal to the last bit.
                    int pos_copy = kPosInvalid;
                    if (av) {
                        pos_copy = 5 * rank;
                        if (plane >= 0) {
                            const int pp = 5 * sm.c_rank[plane] + 1 + last;
                            pos_copy = pp < pos_copy ? pp : pos_copy;
                        }
                    }
                    bool near = false;
                    if (__any_sync(kFull, tied)) {
                        // copies that agree in their high words: exact order of the copies
                        sm.key[li] = kcopy;
                        sm.pos[li] = (uint16_t)pos_copy;
                        __syncwarp();
                        if (tied && av) {
                            const int r = exact_rank<GroupSmem<G, LM, PT>, COUNT>(sm, G, li, kcopy, pos_copy);
                            base = r & 0xff;
                            near = near || (r >> 16);
                        }
                        __syncwarp();
                    }
                    ranks_changed = run && (full || !order_ok);

                    // ---- the extensions that compete, listed in dict order (parent rank, symbol) is not
                    // needed: their positions travel with them
                    int n_ext = 0;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const double ec = c == 0 ? e0 : c == 1 ? e1 : c == 2 ? e2 : e3;
                        const bool comp = full && (c == 0 ? x0 : c == 1 ? x1 : c == 2 ? x2 : x3) >= tau;
                        const unsigned bal = GBALLOT(comp);
                        if (comp) {
                            const int idx = G + n_ext + __popc(bal & belowg);
                            RADIAN_ASSERT(idx < 5 * G);
                            sm.key[idx] = (unsigned long long)__double_as_longlong(ec);
                            sm.pos[idx] = (uint16_t)(5 * rank + 1 + c);
                            sm.src[idx] = (uint8_t)(li * 4 + c);
                        }
                        n_ext += __popc(bal);
                    }
                    int nmax = n_ext;
#pragma unroll
                    for (int o = 16; o >= G; o >>= 1) {
                        const int x = __shfl_xor_sync(kFull, nmax, o);
                        nmax = x > nmax ? x : nmax;
                    }
                    __syncwarp();
                    const int m = G + n_ext;
                    int new_rank = av ? base : 255;
                    int n_new = 0;
                    if (nmax <= G) {
                        // The usual case, at most one extension per lane: lane e keeps extension e.
                        // rank of a copy = its rank among the copies (known) + the extensions above it;
                        // rank of an extension = the copies and the extensions above it.  Counted on
                        // the high words, every lane against all candidates; exact whenever the high
                        // words are all distinct, which the sum of the ranks proves (any tie makes
                        // it fall short of mv(mv-1)/2).
                        const bool mine = li < n_ext;
                        const unsigned long long myk = mine ? sm.key[G + li] : 0ull;
                        const int myp = mine ? (int)sm.pos[G + li] : kPosInvalid;
                        const uint32_t myx = (uint32_t)(myk >> 32);
                        sm.k32[li] = kc32;
                        sm.k32[G + li] = myx;  // (zero beyond the list: counts for nothing)
                        __syncwarp();
                        int my_rank = 0, up = 0;
                        const uint4 *kv = reinterpret_cast<const uint4 *>(sm.k32);
#pragma unroll
                        for (int j = 0; j < G / 4; ++j) {
                            const uint4 k4 = kv[j];
                            my_rank += (k4.x > myx) + (k4.y > myx) + (k4.z > myx) + (k4.w > myx);
                        }
                        const int jend = G / 4 + (nmax + 3) / 4;
#pragma unroll 1
                        for (int j = G / 4; j < jend; ++j) {
                            const uint4 k4 = kv[j];
                            my_rank += (k4.x > myx) + (k4.y > myx) + (k4.z > myx) + (k4.w > myx);
                            up += (k4.x > kc32) + (k4.y > kc32) + (k4.z > kc32) + (k4.w > kc32);
                        }
                        if (av) new_rank = base + up;
                        int ssum = (av ? new_rank : 0) + (mine ? my_rank : 0);
#pragma unroll
                        for (int o = G / 2; o > 0; o >>= 1) ssum += __shfl_xor_sync(kFull, ssum, o);
                        const int mv = na + n_ext;
                        const bool xtie = run && ssum != mv * (mv - 1) / 2;
                        if (__any_sync(kFull, xtie)) {
                            // two candidates agree in their high words: copies and extensions against the
                            // whole list, exactly (float64 bits desc, position asc)
                            sm.key[li] = kcopy;
                            sm.pos[li] = (uint16_t)pos_copy;
                            __syncwarp();
                            if (xtie && av) {
                                const int r = exact_rank<GroupSmem<G, LM, PT>, COUNT>(sm, m, li, kcopy, pos_copy);
                                new_rank = r & 0xff;
                                near = near || (r >> 16);
                            }
                            if (xtie && mine) {
                                const int r = exact_rank<GroupSmem<G, LM, PT>, COUNT>(sm, m, G + li, myk, myp);
                                my_rank = r & 0xff;
                                near = near || (r >> 16);
                            }
                            __syncwarp();
                        }
                        const bool isnew = run && mine && my_rank < bw;
                        const unsigned nbal = GBALLOT(isnew);
                        if (mine) sm.rnk[G + li] = (uint8_t)my_rank;
                        RADIAN_ASSERT(__popc(nbal) <= G);
                        if (isnew) sm.newlist[__popc(nbal & belowg)] = (uint8_t)(G + li);
                        n_new = __popc(nbal);
                    } else {
                        // More extensions than lanes (a beam that is still filling up): every candidate
                        // is ranked against every other one, exactly
                        sm.key[li] = kcopy;
                        sm.pos[li] = (uint16_t)pos_copy;
                        __syncwarp();
                        if (run) near = exact_rank_all<G, GroupSmem<G, LM, PT>, COUNT>(sm, m, li) || near;
                        __syncwarp();
                        if (run) new_rank = av ? (int)sm.rnk[li] : 255;
                        for (int b0 = G; b0 < G + nmax; b0 += G) {
                            const int idx = b0 + li;
                            const bool isnew = run && idx < m && sm.rnk[idx] < bw;
                            const unsigned bal = GBALLOT(isnew);
                            RADIAN_ASSERT(!isnew || n_new + __popc(bal & belowg) < G);
                            if (isnew) sm.newlist[n_new + __popc(bal & belowg)] = (uint8_t)idx;
                            n_new += __popc(bal);
                        }
                    }
                    if (COUNT) n_tie += (GBALLOT(near) != 0u);

                    const bool survive = av && new_rank < bw;
                    const unsigned evb = GBALLOT(av && !survive);
                    const unsigned survb = GBALLOT(survive);
                    const unsigned freeb = GBITS & ~survb;

                    if (__any_sync(kFull, n_new > 0)) {
                        __syncwarp();
                        const int ford = __popc(freeb & belowg);
                        const bool take = run && !survive && ford < n_new;
                        const int item = take ? (int)sm.newlist[ford] : 0;
                        const int s = take ? (int)sm.src[item] : li * 4;
                        const int ls = (s >> 2) + gshift;
                        const int c = s & 3;
                        // parent state, read before anybody overwrites it
                        const int lp = ls - gshift;
                        const uint32_t p_ctx = sm.c_ctx[lp];
                        const int p_len = sm.c_len[lp];
                        const int p_node = sm.c_node[lp];
                        const unsigned long long p_h = sm.c_h[lp];
                        const int p_last = __shfl_sync(kFull, last, ls);
                        double p_r = 0.0;
                        bool p_g = false;
                        int keyctx = -1;  // extend-context of my new beam if the model does not hold it
                        if (LM) {
                            p_g = __shfl_sync(kFull, (int)gext, ls) != 0;
                            // every lane waits for its own row first; the parent's row is then complete
                            cp_async_wait_all();
                            __syncwarp();
                            if (take && p_g) p_r = sm.row[(ls - gshift) * 4 + c];
                        }
                        __syncwarp();  // all reads of parent rows and parent state done before any is replaced
                        int mylen = 0;  // length of the beam this lane holds after the frame (orphans and new beams)
                        unsigned long long myh = 0;
                        if (survive) {
                            ptot = nptot;
                            pnb = npnb;
                            pb = npb;
                            rank = new_rank;
                            if (plane >= 0 && ((evb >> plane) & 1u)) plane = -1;
                        } else if (take) {
                            const double sc = __longlong_as_double((long long)sm.key[item]);
                            ptot = sc;
                            pnb = sc;
                            pb = 0.0;
                            rank = (int)sm.rnk[item];
                            const int node = top + ford;
                            const int len = p_len + 1;
                            const uint32_t ctx = (p_ctx << 2) | (uint32_t)c;
                            last = c;
                            myh = hash_step(p_h, c);
                            mylen = len;
                            sm.c_node[li] = node;
                            sm.c_len[li] = len;
                            sm.c_ctx[li] = ctx;
                            sm.c_hp[li] = p_h;
                            sm.c_h[li] = myh;
                            plane = ((survb >> (ls - gshift)) & 1u) ? (ls - gshift) : -1;
                            prep = (c == p_last) ? 1 : 0;
                            alive = true;
                            RADIAN_ASSERT(node > 0 && node < cap && p_node >= 0 && p_node < node && item >= G && item < 5 * G);
                            arena[node] = ((uint32_t)p_node << 2) | (uint32_t)c;
                            if (LM) {
                                gcopy = p_g;
                                rcopy = p_r;
                                gext = false;
                                if (len >= L) {
                                    const uint32_t ci = ctx & ctx_mask;
                                    const uint32_t gwd = __ldg(a.gate + (ci >> 5));
                                    if (a.miss != nullptr && ((__ldg(a.miss + (ci >> 5)) >> (ci & 31u)) & 1u)) keyctx = (int)ci;
                                    const double *row = a.table + (size_t)ci * 4;
                                    cp_async<16>(&sm.row[li * 4], row);
                                    cp_async<16>(&sm.row[li * 4 + 2], row + 2);
                                    gext = (gwd >> (ci & 31u)) & 1u;
                                }
                            }
                        } else if (run) {
                            alive = false;
                            ptot = pnb = pb = 0.0;
                        }
                        if (run) {
                            top += n_new;
                            na = __popc(survb) + n_new;
                        }
                        if (LM && a.miss != nullptr && __any_sync(kFull, keyctx >= 0)) {
                            // A kept beam whose extend-context the model does not hold: the reference
                            // raises KeyError at lm[context] (decode.py:83) when it processes the beam
                            // in the next frame, best rank first; after the last frame it never looks.
                            int rk = (keyctx >= 0 && tb + it0 + it + 1 < T) ? rank : 0x7fff;
                            int rmin = rk;
#pragma unroll
                            for (int o = G / 2; o > 0; o >>= 1) {
                                const int x = __shfl_xor_sync(kFull, rmin, o);
                                rmin = x < rmin ? x : rmin;
                            }
                            if (run && rmin != 0x7fff) {
                                if (rk == rmin) a.out_len[read] = keyctx;  // the context index, for the message
                                GIVE_UP(RADIAN_READ_KEY_ERROR);
                            }
                        }
                        // A surviving beam whose parent labeling was just (re)created points at it again.
                        // One 64-bit match over the warp pairs them: a new beam offers hash + (length
                        // + 1) x K, an orphan asks for parent hash + length x K, everybody else holds a
                        // value of its own.
                        {
                            bool orphan = survive && plane < 0;
                            if (orphan) mylen = sm.c_len[li];
                            orphan = orphan && mylen > 0;
                            if (__any_sync(kFull, orphan)) {
                                constexpr unsigned long long kLenMix = 0xD6E8FEB86659FD93ull;
                                unsigned long long mv = 0x8000000000000000ull + (unsigned)lane;
                                if (orphan) mv = sm.c_hp[li] + (unsigned long long)mylen * kLenMix;
                                if (take) mv = myh + (unsigned long long)(mylen + 1) * kLenMix;
                                const unsigned same = __match_any_sync(kFull, mv);
                                const unsigned cand = same & (GBALLOT(take) << gshift);
                                const int pl = cand ? __ffs(cand) - 1 : lane;  // (at most one new beam per labeling)
                                const int z_last = __shfl_sync(kFull, last, pl);
                                if (orphan && cand) {
                                    plane = pl - gshift;
                                    prep = (z_last == last) ? 1 : 0;
                                }
                            }
                        }
                        // the beam set changed: refresh which extensions are merged into a live child
                        sm.kill[li] = 0u;
                        __syncwarp();
                        if (run && alive && plane >= 0) reinterpret_cast<uint8_t *>(sm.kill)[plane * 4 + last] = 0x80;
                        __syncwarp();
                        if (run) km = alive ? (0x80808080u & ~sm.kill[li]) : 0u;
                        if (LM && run && alive && gext) {
                            // table part of the quiet-frame bound, over the symbols that are still
                            // candidates of their own.  The row of a beam created in this frame is in
                            // flight: until the next frame that comes this way it is bounded by the
                            // largest entry of the whole table (a.rcap), so nobody waits for a gather.
                            rmax = take ? a.rcap : row_bound(&sm.row[li * 4], km);
                            rmax_prov = take;
                        }
                    } else if (survive) {
                        ptot = nptot;
                        pnb = npnb;
                        pb = npb;
                        rank = new_rank;
                    }
                }
                if (__any_sync(kFull, ranks_changed)) {
                    // successor lane of every beam and the lane of the best one
                    __syncwarp();
                    RADIAN_ASSERT(!(run && alive) || (rank >= 0 && rank < na && na <= G));
                    if (run && alive) sm.newlist[rank] = (uint8_t)li;
                    sm.c_rank[li] = rank;
                    __syncwarp();
                    if (run) {
                        succ = (alive && rank + 1 < na) ? (int)sm.newlist[rank + 1] + gshift : lane;
                        first_lane = (int)sm.newlist[0] + gshift;
                        last_lane = (int)sm.newlist[na - 1] + gshift;
                    }
                    // should my successor ever have exactly my score: is it rightly behind me?  (dict
                    // insertion position of the next frame's copies, from the new ranks)
                    int npos = 5 * rank;
                    if (run && alive && plane >= 0) {
                        const int pp = 5 * sm.c_rank[plane] + 1 + last;
                        npos = pp < npos ? pp : npos;
                    }
                    const int spos = __shfl_sync(kFull, npos, succ);
                    if (run) tie_ok = npos < spos;
                    __syncwarp();
                }
                {
                    const unsigned long long mine = (unsigned long long)__double_as_longlong(ptot);
                    const unsigned long long next = __shfl_sync(kFull, mine, succ);
                    hitie = run && alive && succ != lane && (mine >> 32) == (next >> 32) &&
                            (mine > next || (mine == next && tie_ok));
                }
                REFRESH();
                ++it;
            }
        }
        if (live) t += nrun;

        // ------------------------------------------------------------ end of read
        int succ_first = __shfl_sync(kFull, succ, first_lane);  // lane of the second best beam
        int best_lane = first_lane;
        {
            // the best two beams, exactly: if they share their high word, their low words may have crossed
            // since the order was last checked on all 64 bits (hitie)
            const unsigned long long mine = (unsigned long long)__double_as_longlong(ptot);
            const unsigned long long p1 = __shfl_sync(kFull, mine, first_lane), p2 = __shfl_sync(kFull, mine, succ_first);
            const bool t1 = __shfl_sync(kFull, (int)tie_ok, first_lane) != 0;
            if (succ_first != first_lane && (p2 > p1 || (p2 == p1 && !t1))) {
                best_lane = succ_first;
                succ_first = first_lane;
            }
        }
        if (live && t >= T) {
            const long long seq_off = a.seq_offsets[read];
            const long long seq_cap = a.seq_offsets[read + 1] - seq_off;
            if (status == 0 && lane == best_lane) {
                const long long n = sm.c_len[li];
                if (n > seq_cap) status = RADIAN_READ_SEQ_OVERFLOW;
                int c = sm.c_node[li];
                for (long long i = n - 1; i >= 0; --i) {
                    RADIAN_ASSERT(c > 0 && c < cap);
                    const uint32_t w = arena[c] & 0x7fffffffu;
                    if (i < seq_cap) a.out_seq[seq_off + i] = (uint8_t)(w & 3u);
                    c = (int)(w >> 2);
                }
                a.out_len[read] = n;
                a.out_score[2 * read] = final_log_score(ptot, kacc);
                if (succ_first == best_lane) a.out_score[2 * read + 1] = NAN;
                a.out_status[read] = status;
                if (a.out_counters) {
                    a.out_counters[4 * read] = n_lookup;
                    a.out_counters[4 * read + 1] = n_combine;
                    a.out_counters[4 * read + 2] = n_tie;
#ifdef RADIAN_READ_TIMES
                    // (diagnostic build: nanoseconds this read spent in its group instead of the gate count)
                    unsigned long long now_;
                    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now_));
                    n_stage2 = ((now_ - t_read0) << 32) | (n_stage2 & 0xffffffffull);
#endif
                    a.out_counters[4 * read + 3] = n_stage2;
                }
            }
            if (status == 0 && succ_first != best_lane && lane == succ_first)
                a.out_score[2 * read + 1] = final_log_score(ptot, kacc);
            if (status > RADIAN_READ_SEQ_OVERFLOW && li == 0) {
                if (status != RADIAN_READ_KEY_ERROR) a.out_len[read] = 0;  // (KeyError: holds the context index)
                a.out_score[2 * read] = NAN;
                a.out_score[2 * read + 1] = NAN;
                a.out_status[read] = status;
            }
            read = -1;
            // leave no beam behind: an idle group must look "in order, nothing competing"
            alive = false;
            succ = lane;
            km = 0u;
            ptot = pnb = pb = 0.0;
        }
    }
#undef GBALLOT
#undef REFRESH
#undef GIVE_UP
#undef RESCALE_CHECK
}

// ------------------------------------------------------------------------------ host side

// lanes (<= 32) or beam slots (wide kernel) reserved per read
static int group_size(int beam_width)
{
    return beam_width <= 8 ? 8 : beam_width <= 16 ? 16 : beam_width <= 32 ? 32 : beam_width <= 64 ? 64 : 128;
}

template <int G, bool LM, typename PT>
static const void *kernel_ptr(bool count, bool stream)
{
    if (stream)
        return count ? (const void *)decode_kernel<G, LM, PT, true, true> : (const void *)decode_kernel<G, LM, PT, false, true>;
    return count ? (const void *)decode_kernel<G, LM, PT, true, false> : (const void *)decode_kernel<G, LM, PT, false, false>;
}

static const void *pick_kernel(int G, bool lm, bool f64, bool count, bool stream)
{
#define RADIAN_PICK(GG)                                                                        \
    if (G == GG) {                                                                             \
        if (lm) return f64 ? kernel_ptr<GG, true, double>(count, stream) : kernel_ptr<GG, true, float>(count, stream); \
        return f64 ? kernel_ptr<GG, false, double>(count, stream) : kernel_ptr<GG, false, float>(count, stream);       \
    }
    RADIAN_PICK(8)
    RADIAN_PICK(16)
    RADIAN_PICK(32)
#undef RADIAN_PICK
    return nullptr;
}

int decode_nursery() { return kNursery; }

int64_t decode_arena_cap(int beam_width, int64_t max_frames, int64_t arena_nodes)
{
    // Old generation <= beam_width x decoded length (see the header comment); decoded length <= T.
    // Small problems get the exact worst case; large ones get lanes x T/32 nodes (the labelings of a
    // read that differ at an old ambiguous position carry separate chains from there on: up to lanes x
    // decoded length nodes, i.e. lanes x T/32 at 32 frames per base or more) and report
    // RADIAN_READ_TRIE_OVERFLOW otherwise (the caller retries those reads with arena_nodes set).
    const int64_t G = group_size(beam_width);
    const int64_t exact = G * (max_frames + 1) + kNursery + 64;
    if (arena_nodes > 0) return arena_nodes < exact ? arena_nodes + kNursery : exact;
    if (exact <= (1 << 16)) return exact;
    int64_t cap = G * (max_frames / 32 + 64) + kNursery;
    return cap < (1 << 16) ? (1 << 16) : cap;
}

int decode_pick(int device, int beam_width, bool lm, bool f64, bool count, bool stream, DecodeLaunch *out)
{
    DeviceInfo di;
    int rc = device_info(device, &di);
    if (rc) return rc;
    const int G = group_size(beam_width);
    int blocks = 0;
    if (G > 32) {
        // beam widths above 32: one warp per read, several beams per lane (decode_wide.cu)
        const void *k = nullptr;
        size_t smem = 0;
        const int warps = wide_pick(beam_width, lm, f64, count, &k, &smem);
        RADIAN_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RADIAN_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        RADIAN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, k, warps * 32, smem));
        if (blocks < 1) blocks = 1;
        out->kernel = k;
        out->smem = smem;
        out->grid = di.sm_count * blocks;
        out->block = warps * 32;
        out->groups_per_block = warps;
        return 0;
    }
    const void *k = pick_kernel(G, lm, f64, count, stream);
    // the kernel streams its global loads once; give shared memory the whole L1 carve-out
    RADIAN_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    RADIAN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, k, kWarpsPerBlock * 32, 0));
    if (blocks < 1) blocks = 1;
    out->kernel = k;
    out->smem = 0;
    out->grid = di.sm_count * blocks;
    out->block = kWarpsPerBlock * 32;
    out->groups_per_block = kWarpsPerBlock * (32 / G);
    return 0;
}

int decode_max_slots(int device, int beam_width)
{
    int best = 0;
    for (int v = 0; v < 16; ++v) {
        DecodeLaunch dl;
        if (decode_pick(device, beam_width, v & 1, v & 2, v & 4, v & 8, &dl)) return -1;
        int s = dl.grid * dl.groups_per_block;
        best = s > best ? s : best;
    }
    return best;
}

// Launch shape of a resident batch.  Measured on B200 (scripts/ab_occupancy.sh, profiles/r2_history.md):
// a read group needs about 750 cycles per frame when its warp is alone on its scheduler and ~90 more
// for every further warp there, so throughput keeps growing with occupancy (3.1e9 frames/s at one CTA
// per SM, 1.05e10 at five) -- but a batch that does not fill the machine many times over is finished
// when its longest read is, and that read wants the opposite: few warps per scheduler, and no
// warp-mates (a read pays for the frames in which its warp-mates change their beam sets; alone in its
// warp it needs about kAlone of the time).  So:
//   * CTAs per SM: the fewest that still decode the bulk of the batch within the time the longest
//     read needs alone, i.e. slots(w) >= total frames / (kAlone x frames of the longest read);
//   * reads of at least excl_frames = max(kAlone x longest, total / slots) frames are run one per
//     warp (decode_kernel, "EXCLUSIVE reads"); none if that is not below the longest read.
// Batches that fill the machine get the full occupancy and no exclusive reads.
static void plan_launch(int max_ctas, int groups_per_block, int warps_per_block, int sm_count, int n_reads,
                        int64_t max_frames, int64_t total_frames, int *per_sm, int *excl_frames)
{
    static const char *env_w = getenv("RADIAN_CTAS_PER_SM"), *env_x = getenv("RADIAN_EXCL_FRAMES"),
                      *env_a = getenv("RADIAN_TUNE_ALONE");
    *per_sm = max_ctas;
    *excl_frames = 0;
    const int gpw = groups_per_block / warps_per_block;
    if (total_frames <= 0) total_frames = (int64_t)n_reads * max_frames;  // unknown: all reads as long as the longest
    const double alone = env_a ? atof(env_a) : (gpw > 1 ? 0.6 : 1.0);  // measured: 234 vs 384 ns per frame, four reads per warp
    if (max_frames > 0 && n_reads > 0) {
        const double need_slots = (double)total_frames / (alone * (double)max_frames);
        int w = (int)(need_slots / ((double)sm_count * groups_per_block)) + 1;
        *per_sm = w < 1 ? 1 : w > max_ctas ? max_ctas : w;
        const double slots = (double)sm_count * *per_sm * groups_per_block;
        const double x = std::max(alone * (double)max_frames, (double)total_frames / slots);
        if (gpw > 1 && x < (double)max_frames) *excl_frames = (int)std::max(x, 1.0);
    }
    if (env_w && atoi(env_w) > 0) *per_sm = atoi(env_w) < max_ctas ? atoi(env_w) : max_ctas;
    if (env_x) *excl_frames = atoi(env_x);
}

int decode_launch(const DecodeArgs &a, bool f64, int device, cudaStream_t stream)
{
    const bool lm = a.table != nullptr;
    DecodeLaunch dl;
    int rc = decode_pick(device, a.beam_width, lm, f64, a.out_counters != nullptr, a.ready != nullptr, &dl);
    if (rc) return rc;
    DeviceInfo di;
    rc = device_info(device, &di);
    if (rc) return rc;
    int per_sm = dl.grid / di.sm_count, excl_frames = 0;
    if (a.beam_width <= 32 && a.ready == nullptr)  // (streamed batches arrive over time: keep every slot)
        plan_launch(per_sm, dl.groups_per_block, dl.block / 32, di.sm_count, a.n_reads, a.max_frames, a.total_frames,
                    &per_sm, &excl_frames);
    // no more groups than reads: extra CTAs would only touch the queue (with exclusive reads a warp
    // may take a single read, and how many reads are that long is only known on the device)
    const int reads_per_block = excl_frames > 0 ? dl.block / 32 : dl.groups_per_block;
    int64_t need = ((int64_t)a.n_reads + reads_per_block - 1) / reads_per_block;
    const int64_t cap = (int64_t)di.sm_count * per_sm;
    int grid = (int)(need < cap ? need : cap);
    if (grid < 1) grid = 1;
    size_t smem = dl.smem;
    if (per_sm < dl.grid / di.sm_count && need >= cap) {
        // fewer CTAs per SM than would fit: pad the request with unused dynamic shared memory so that
        // exactly per_sm of them fit an SM and the hardware spreads the grid evenly
        cudaFuncAttributes fa;
        RADIAN_CUDA(cudaFuncGetAttributes(&fa, dl.kernel));
        const size_t per_cta = (size_t)di.smem_per_sm / per_sm;
        const size_t fixed = fa.sharedSizeBytes + 1024;  // static + the 1 KB the system reserves per CTA
        size_t dyn = per_cta > fixed ? ((per_cta - fixed) & ~(size_t)127) : 0;
        if (dyn + fa.sharedSizeBytes > (size_t)di.max_smem_optin) dyn = (size_t)di.max_smem_optin - fa.sharedSizeBytes;
        RADIAN_CUDA(cudaFuncSetAttribute(dl.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
        smem = dyn;
    }
    DecodeArgs args = a;
    args.excl_frames = excl_frames;
    void *params[] = {(void *)&args};
    RADIAN_CUDA(cudaLaunchKernel(dl.kernel, dim3(grid), dim3(dl.block), params, smem, stream));
    return 0;
}

}  // namespace radian
