// CTC prefix beam search for beam widths above 32 (33..128), sm_100a.
//
// Same computation as decode.cu (radian/decode.py:100-212 of the reference: frame loop 141-204,
// COPY 150-175, EXTEND 177-201, apply_rna_model 79-96, combine_dists 52-64, final pick 207-210)
// and the same building blocks -- linear-domain float64 scores rescaled by exact powers of two,
// frame records shared through shared memory, RNA rows gathered once per new beam, copy/extend
// merge through a parent pointer, generational back-pointer arena -- laid out differently:
//   * one warp owns a read; beam b lives in slot b/32 of lane b%32, so a lane carries BPL = 2 or 4
//     beams in registers;
//   * quiet frames (nine in ten: order intact, beam full, no extension can reach the worst copy by the
//     integer bound on the high words) run in a loop of their own that only updates the scores:
//     the parent's old score and the successor's new high word go through shared memory, one vote
//     per frame; what the loop needs of a beam besides its scores is refreshed after every other frame;
//   * every other frame computes the extensions; those that can reach the beam are ranked with the
//     copies: (float64 bit pattern desc, dict insertion position asc), the reference's stable sort over
//     its insertion-ordered dict (decode.py:35-39, 145) -- on high words by bisection in the copies'
//     descending order plus a histogram, on all 64 bits wherever two high words agree;
//   * shared memory per read decides how many reads an SM holds (15 at BPL 2, 8 at BPL 4): one-warp
//     CTAs, compact records, staging arrays aliased onto arrays that are idle at that point.
// Where the cycles of a read go: -DRADIAN_WIDE_PROBE, scripts/wide_probe.py, DESIGN.md 4.2.
#include "decode_common.cuh"

namespace radian {

// -DRADIAN_WIDE_PROBE (diagnostic build, scripts/wide_probe.py): a lap timer per warp; RADIAN_LAP(ch)
// charges the cycles since the previous mark to channel ch, totals over all reads in g_lap.
#ifdef RADIAN_WIDE_PROBE
__device__ unsigned long long g_lap[24];
#define RADIAN_LAP(ch)                                                              \
    do {                                                                            \
        const long long now_ = clock64();                                           \
        if (lane == 0) sm.lap[ch] += (unsigned long long)(now_ - lap_t);            \
        lap_t = now_;                                                               \
    } while (0)
#else
#define RADIAN_LAP(ch)
#endif

constexpr int kWideWarps = 1;  // warps (reads) per CTA: shared memory per read decides how many fit an SM,
                               // and one-warp CTAs waste the least of it

template <int BPL, bool LM, typename PT>
struct __align__(16) WideSmem {
    static constexpr int NB = 32 * BPL;
    static constexpr int REC = LM ? 14 : 6;   // doubles per frame, compact record of decode_common.cuh
    double rec[32 * REC];
    double row[LM ? NB * 4 : 4];      // RNA table row of every beam's extend-context (cp.async target)
    double ex[NB * 2];                // {pr_total, pr_blank} of every beam before the frame
    double zero[2];                   // what a beam without a live parent merges with
#ifdef RADIAN_WIDE_PROBE
    unsigned long long lap[24];
#endif
    unsigned long long key[5 * NB];   // candidate scores: [0,NB) copies by beam id, then extensions
    uint32_t k32[5 * NB];             // high words of the candidate scores (incremental ranking)
    PT raw[32 * 5];                   // next tile of posterior rows, landed by cp.async
    uint32_t sctx[NB];                // staged: packed context
    int32_t slen[NB];                 // staged: labeling length
    uint32_t kill[NB];                // byte c of word b: extension (b,c) merged into a live child's copy
    uint16_t pos[5 * NB];             // dict insertion position of the candidate
    // Two staging arrays of the creation step share storage with arrays that are idle by then (shared
    // memory per read decides how many reads an SM holds): the labeling hashes use the extensions' part
    // of pos (read last by the exact ranking, rewritten by the next frame's candidate list), the
    // arena nodes use kill (histogram of the ranking before, zeroed and rebuilt after the creation).
    __device__ unsigned long long *sh() { return reinterpret_cast<unsigned long long *>(&pos[NB]); }
    __device__ int32_t *snode() { return reinterpret_cast<int32_t *>(kill); }
    uint16_t src[5 * NB];             // beam*4+c of an extension candidate
    uint16_t rnk[5 * NB];             // rank of the candidate
    uint16_t newlist[NB];             // candidate indices of the new beams
    uint16_t byrank[NB];              // beam id by rank
    uint16_t srank[NB];               // staged: rank of every beam
    uint8_t sgext[NB];                // staged: gate bit of every beam's extend-context
    uint8_t slast[NB];                // staged: last symbol of every beam
};

// What the launch shape rests on (228 KB of shared memory per SM, 1 KB of it reserved per CTA): fifteen
// one-warp CTAs per SM at BPL 2, eight at BPL 4 (float32 posteriors, model on).
#ifndef RADIAN_WIDE_PROBE
static_assert(15 * (sizeof(WideSmem<2, true, float>) + 1024) <= 233472, "BPL 2: fifteen reads per SM");
static_assert(8 * (sizeof(WideSmem<4, true, float>) + 1024) <= 233472, "BPL 4: eight reads per SM");
#endif

constexpr unsigned long long kHashEmpty = 0x243F6A8885A308D3ull;
constexpr int kNoRmaxW = (int)0x80000000;

template <int BPL, bool LM, typename PT, bool COUNT>
__global__ void __launch_bounds__(kWideWarps * 32, BPL == 2 ? 15 : 8)
decode_wide_kernel(const DecodeArgs a)
{
    using SM = WideSmem<BPL, LM, PT>;
    static_assert(offsetof(SM, pos) % 8 == 0 && (SM::NB * 2) % 8 == 0, "sh() must be 8-byte aligned");
    constexpr int NB = SM::NB;
    constexpr int REC = SM::REC;
    extern __shared__ __align__(16) unsigned char wide_smem_raw[];
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    SM &sm = reinterpret_cast<SM *>(wide_smem_raw)[warp];
    const unsigned below = (1u << lane) - 1u;
    const int slot = blockIdx.x * kWideWarps + warp;

    const int bw = a.beam_width;
    const int L = a.L;
    const uint32_t ctx_mask = LM ? (uint32_t)((1ull << (2 * L)) - 1ull) : 0u;
    const int cap = a.arena_cap;
    uint32_t *const arena = a.arena + (size_t)slot * (size_t)(cap + kNursery);
    uint32_t *const fwd = arena + cap;

    // ---- per-beam state, slot s of this lane is beam s*32+lane; a dead beam keeps zero scores
    double ptot[BPL], pnb[BPL], pb[BPL], rcopy[BPL];
    unsigned long long hp[BPL];  // hash of the parent labeling
    uint32_t ctx[BPL], km[BPL];
    int len[BPL], node[BPL], rank[BPL], plane[BPL], last[BPL], succ[BPL];
    bool alive[BPL], gext[BPL], gcopy[BPL];
    bool tie_ok[BPL];  // a successor with exactly my score is still in the right place (later insertion position)
    int prep[BPL];  // 1 if the live parent (plane) ends in the same symbol as this beam
    int rmax[BPL];  // max high word of the unmerged entries of this beam's table row, or kNoRmaxW
    bool rprov[BPL];  // this beam's table row is still in flight (its bound is the table-wide one, a.rcap)
    // what a quiet frame needs besides the three scores, set by REFRESH after every frame that is not
    // quiet (see decode.cu): where the copy emission of the beam is in a record (x, y), the score of
    // the parent whose extension merges into it (or a zero), the high word of its successor, and
    // whether that one has to be strictly smaller (q_inc = 1) or, for a pair whose order the long way
    // has confirmed on all 64 bits although the high words agree, may be equal (0)
    unsigned q_xa[BPL], q_ya[BPL], q_pa[BPL], q_ks[BPL];
    uint32_t q_inc[BPL];
    const unsigned a_rec = (unsigned)__cvta_generic_to_shared(&sm.rec[0]);
    const unsigned a_ex = (unsigned)__cvta_generic_to_shared(&sm.ex[0]);
    const unsigned a_k32 = (unsigned)__cvta_generic_to_shared(&sm.k32[0]);
    const unsigned a_zero = (unsigned)__cvta_generic_to_shared(&sm.zero[0]);
    if (lane < 2) sm.zero[lane] = 0.0;
    unsigned c_lookup = 0, c_combine = 0;  // COUNT: table lookups / combined emissions of one quiet frame

    while (true) {
        // ------------------------------------------------------------ fetch a read
        int idx = 0;
        if (lane == 0) {
            idx = atomicAdd(a.queue, 1);
            if (a.ready != nullptr && idx < a.n_reads) {
                int seen = -2;
                unsigned t0 = 0;
                while (true) {
                    const unsigned long long v = *(const volatile unsigned long long *)a.ready;
                    const int r0 = (int)(unsigned)v, r1 = (int)(unsigned)(v >> 32);
                    const int landed = r0 < r1 ? r0 : r1;
                    if (landed > idx) break;
                    // stalled-transfer guard, see decode.cu
                    const unsigned now = timer_units();
                    if (*(const volatile int *)(a.queue + 1) != 0) {
                        idx = 0x7fffffff;
                        break;
                    }
                    if (landed != seen) {
                        seen = landed;
                        t0 = now;
                    } else if (now - t0 > kStallUnits) {
                        atomicExch(a.queue + 1, 1);
                        idx = 0x7fffffff;
                        break;
                    }
                    __nanosleep(400);
                }
                __threadfence();
            }
        }
        idx = __shfl_sync(kFull, idx, 0);
        if (idx >= a.n_reads) break;
        const int read = a.order ? a.order[idx] : idx;
        const long long foff = a.frame_offsets[read];
        const int T = (int)(a.frame_offsets[read + 1] - foff);
        const PT *rp = (const PT *)a.post + foff * 5;
#pragma unroll
        for (int s = 0; s < BPL; ++s) {
            // initial beam: the empty labeling, pr_blank = pr_total = log 1 (decode.py:128-132)
            alive[s] = (s == 0 && lane == 0);
            ptot[s] = alive[s] ? 1.0 : 0.0;
            pb[s] = ptot[s];
            pnb[s] = 0.0;
            rcopy[s] = 0.0;
            hp[s] = 0;
            ctx[s] = 0;
            len[s] = 0;
            node[s] = 0;
            rank[s] = 0;
            plane[s] = -1;
            last[s] = 0;
            succ[s] = s * 32 + lane;
            tie_ok[s] = false;
            km[s] = alive[s] ? 0x80808080u : 0u;
            gext[s] = gcopy[s] = false;
            prep[s] = 0;
            rmax[s] = kNoRmaxW;
            rprov[s] = false;
        }
        int first = 0, last_b = 0;  // beam ids of the best and the worst ranked beam
        int top = 1, old_top = 1, na = 1, status = 0;  // node 0 = the empty labeling
        long long kacc = 0;
        unsigned long long n_lookup = 0, n_combine = 0, n_tie = 0, n_diag = 0;  // n_diag: see include/radian_b200.h
#ifdef RADIAN_WIDE_PROBE
        if (lane < 24) sm.lap[lane] = 0;
        long long lap_t = clock64();
#endif

        __syncwarp();
        if (lane < T) prefetch_row(&sm.raw[lane * 5], rp, lane);

        // RESCALE by an exact power of two when the best beam has left [2^300, 2^900): back to 2^600.
        // Checked once per tile and before every frame that is not quiet; in between, the quiet test
        // refuses a frame in which the worst kept beam is not a normal number, and the long way
        // reports RADIAN_READ_RANGE for a beam that underflows with non-zero factors: nothing is lost
        // silently (see decode.cu).
        auto rescale_check = [&]() {
            int hi = 0;
#pragma unroll
            for (int s = 0; s < BPL; ++s) {
                const int x = __shfl_sync(kFull, __double2hiint(ptot[s]), first & 31);
                if ((first >> 5) == s) hi = x;
            }
            const int exb = (hi >> 20) & 0x7ff;
            if ((unsigned)(exb - (1023 + 300)) >= 600u) {
                const int k1 = exb == 0 ? 1000 : 1023 - exb;  // a subnormal best first comes up by 2^1000
                const double s1 = __hiloint2double((1023 + k1) << 20, 0);
                const double s2 = __hiloint2double((1023 + 600) << 20, 0);
#pragma unroll
                for (int s = 0; s < BPL; ++s) {
                    ptot[s] = __dmul_rn(__dmul_rn(ptot[s], s1), s2);
                    pnb[s] = __dmul_rn(__dmul_rn(pnb[s], s1), s2);
                    pb[s] = __dmul_rn(__dmul_rn(pb[s], s1), s2);
                }
                kacc -= k1 + 600;
            }
        };
        // REFRESH: what the quiet frames use of a beam, after anything about the beams has changed
        auto refresh = [&]() {
            __syncwarp();
#pragma unroll
            for (int s = 0; s < BPL; ++s)
                sm.key[s * 32 + lane] = (unsigned long long)__double_as_longlong(ptot[s]);
            __syncwarp();
            unsigned cl = 0, cc = 0;
#pragma unroll
            for (int s = 0; s < BPL; ++s) {
                const int b = s * 32 + lane;
                const bool gc = LM && gcopy[s];
                q_xa[s] = a_rec + (unsigned)((gc ? 6 + last[s] : last[s]) * 8);
                q_ya[s] = a_rec + (unsigned)((gc ? 11 : 10) * 8);
                q_pa[s] = (alive[s] && plane[s] >= 0) ? a_ex + (unsigned)((plane[s] * 2 + prep[s]) * 8) : a_zero;
                q_ks[s] = a_k32 + (unsigned)(succ[s] * 4);
                const bool has = alive[s] && succ[s] != b;
                const unsigned long long mine = (unsigned long long)__double_as_longlong(ptot[s]);
                const unsigned long long next = sm.key[succ[s]];
                const bool hitie = has && (mine >> 32) == (next >> 32) && (mine > next || (mine == next && tie_ok[s]));
                q_inc[s] = (has && !hitie) ? 1u : 0u;
                if (LM && gext[s] && rmax[s] == kNoRmaxW)
                    rmax[s] = rprov[s] ? a.rcap : row_bound(&sm.row[b * 4], km[s]);
                if (COUNT && LM) {
                    const bool lm_copy = alive[s] && len[s] >= L + 1;  // decode.py:157
                    const bool lm_ext = alive[s] && len[s] >= L;       // decode.py:180
                    cl += __popc(__ballot_sync(kFull, lm_copy)) + __popc(__ballot_sync(kFull, lm_ext));
                    cc += __popc(__ballot_sync(kFull, lm_copy && gcopy[s])) + __popc(__ballot_sync(kFull, lm_ext && gext[s]));
                }
            }
            c_lookup = cl;
            c_combine = cc;
            __syncwarp();
        };
        refresh();

        int t = 0;
        while (t < T && status == 0) {
            // -------------------------------------------------------- tile refill
            if ((t & 31) == 0) {
                cp_async_wait_all();
                __syncwarp();
                int kf = 0;
                if (t + lane < T) kf = make_record<LM, true, sizeof(PT) == 8, true>(&sm.raw[lane * 5], a.s_thr, &sm.rec[lane * REC]);
                __syncwarp();
                if (t + 32 + lane < T) prefetch_row(&sm.raw[lane * 5], rp, t + 32 + lane);
                if (sizeof(PT) == 8 && __any_sync(kFull, kf != 0)) {
                    // exponents taken out of tiny rows (make_record): stored score = true score x 2^-kacc
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) kf += __shfl_xor_sync(kFull, kf, o);
                    kacc -= kf;
                }
                rescale_check();
            }

            // -------------------------------------------------------- QUIET frames
            // The common case (decode.cu): the order of the copies holds, the beam is full and, by the
            // integer bound on the high words (largest symbol of the frame, table bound of the beam),
            // no extension can reach the worst copy.  Such a frame is three score updates per beam,
            // two shared-memory exchanges (the parent's old score, the successor's new high word) and
            // one vote; nothing else about the beams is touched.
            RADIAN_LAP(13);
            if (na >= bw) {
                int it = t & 31;
                const int tile0 = t - it;
                const int nend = (T - tile0) < 32 ? (T - tile0) : 32;
                const unsigned a_kw = a_k32 + (unsigned)(last_b * 4);
#pragma unroll 1
                for (; it < nend; ++it) {
                    const unsigned ro = (unsigned)it * (unsigned)(REC * 8);
                    const double P4 = lds_f64(a_rec + ro + 32);
                    double g_ = 0.0;
                    int4 gi = make_int4(0, 0, 0, 0);
                    if (LM) {
                        g_ = lds_f64(a_rec + ro + 40);
                        gi = lds_i4(a_rec + ro + 96);   // {gate, hS, zP, zQ}, decode_common.cuh
                    } else {
                        gi.z = lds_i32(a_rec + ro + 40);
                    }
                    double nptot[BPL], npnb[BPL], npb[BPL], dl[BPL];
#pragma unroll
                    for (int s = 0; s < BPL; ++s) {
                        // COPY (decode.py:150-175): (rcopy * gate + x) * y is P[last], or with the gate open
                        // and a gated copy-context (r + p/S) * (S/2)
                        dl[s] = lds_f64(q_xa[s] + ro);
                        if (LM) dl[s] = __dmul_rn(__dadd_rn(__dmul_rn(rcopy[s], g_), dl[s]), lds_f64(q_ya[s] + ro));
                        npnb[s] = __dmul_rn(pnb[s], dl[s]);
                        npb[s] = __dmul_rn(ptot[s], P4);
                        nptot[s] = __dadd_rn(npb[s], npnb[s]);
                        asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a_ex + (unsigned)((s * 32 + lane) * 16)), "d"(ptot[s]),
                                     "d"(pb[s])
                                     : "memory");
                    }
                    __syncwarp();
                    uint32_t kc[BPL];
                    double pv[BPL];
#pragma unroll
                    for (int s = 0; s < BPL; ++s) pv[s] = lds_f64_volatile(q_pa[s]);
#pragma unroll
                    for (int s = 0; s < BPL; ++s) {
                        // MERGE with the parent's extension by my last symbol
                        const double v = __dmul_rn(pv[s], dl[s]);
                        npnb[s] = __dadd_rn(npnb[s], v);
                        nptot[s] = __dadd_rn(nptot[s], v);
                        kc[s] = (uint32_t)__double2hiint(nptot[s]);
                        asm volatile("st.shared.u32 [%0], %1;" ::"r"(a_k32 + (unsigned)((s * 32 + lane) * 4)), "r"(kc[s]) : "memory");
                    }
                    __syncwarp();
                    uint32_t kworst;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(kworst) : "r"(a_kw) : "memory");
                    bool quiet = kworst >= 0x00100000u;
#pragma unroll
                    for (int s = 0; s < BPL; ++s) {
                        uint32_t ksucc;
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(ksucc) : "r"(q_ks[s]) : "memory");
                        const int z = (LM && gext[s]) ? max(gi.w, rmax[s]) + gi.y : gi.z;
                        quiet = quiet && kc[s] >= ksucc + q_inc[s] && __double2hiint(ptot[s]) + z < (int)kworst;
                    }
                    if (!__all_sync(kFull, quiet)) break;
                    if (COUNT && LM) {
                        n_lookup += c_lookup;
                        if (gi.x != 0) {
                            n_combine += c_combine;
                            n_diag += 1ull << 32;
                        }
                    }
#pragma unroll
                    for (int s = 0; s < BPL; ++s) {
                        ptot[s] = nptot[s];  // (a dead beam's new values are zero as well)
                        pnb[s] = npnb[s];
                        pb[s] = npb[s];
                    }
                }
                t = tile0 + it;
                RADIAN_LAP(0);
                if (it == nend) continue;
            }
            bool need_ = false;
            do {  // one frame the long way (a `break` leaves it with `status` set)
            if (LM) {
                // table bounds that were provisional (the row was in flight): the real ones now
                bool pv = false;
#pragma unroll
                for (int s = 0; s < BPL; ++s) pv = pv || (gext[s] && rprov[s]);
                if (__any_sync(kFull, pv)) {
                    cp_async_wait_all();
#pragma unroll
                    for (int s = 0; s < BPL; ++s)
                        if (gext[s] && rprov[s]) {
                            rprov[s] = false;
                            rmax[s] = kNoRmaxW;
                        }
                }
            }

            RADIAN_LAP(1);
            // -------------------------------------------------------- nursery collection
            if (top + NB > old_top + kNursery || top + NB > cap) {
                // 1. mark the nursery nodes reachable from a live beam
#pragma unroll
                for (int s = 0; s < BPL; ++s) {
                    int cur = node[s];
                    bool walking = alive[s] && cur >= old_top;
                    while (walking) {
                        const uint32_t w = arena[cur];
                        if (w >> 31) break;
                        arena[cur] = w | 0x80000000u;
                        cur = (int)(w >> 2);
                        walking = cur >= old_top;
                    }
                }
                __syncwarp();
                // 2. slide marked nodes down in index order (parents precede children); fwd[] keeps
                //    the new index of every moved node for its children and for the beams
                int cnt = old_top;
                for (int base = old_top; base < top; base += 32) {
                    const int i = base + lane;
                    const uint32_t w = (i < top) ? arena[i] : 0u;
                    const bool mk = (w >> 31) != 0;
                    const unsigned bal = __ballot_sync(kFull, mk);
                    const int ni = cnt + __popc(bal & below);
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
                        arena[ni] = ((uint32_t)npar << 2) | (w & 3u);
                        fwd[i - old_top] = (uint32_t)ni;
                    }
                    cnt += __popc(bal);
                    __syncwarp();
                }
#pragma unroll
                for (int s = 0; s < BPL; ++s)
                    if (alive[s] && node[s] >= old_top) node[s] = (int)fwd[node[s] - old_top];
                __syncwarp();
                old_top = cnt;
                top = cnt;
                if (top + NB > cap) {
                    status = RADIAN_READ_TRIE_OVERFLOW;  // reported; remaining frames are skipped
                    break;
                }
            }

            rescale_check();

            RADIAN_LAP(2);
            // -------------------------------------------------------- one frame
            // (the structure of decode.cu's frame: copy, child-side merge, integer quiet test; the
            // extension scores are only computed when some extension may reach the beam)
            const double *rec = &sm.rec[(t & 31) * REC];
            const int *reci = reinterpret_cast<const int *>(rec);
            const double P4 = rec[4];
            bool fgate = false;
            int hS = 0;
            if (LM) {
                const int2 gs = *reinterpret_cast<const int2 *>(reci + 24);
                fgate = gs.x != 0;
                hS = gs.y;
            }
            double nptot[BPL], npnb[BPL], npb[BPL], dl[BPL];
#pragma unroll
            for (int s = 0; s < BPL; ++s) {
                const int b = s * 32 + lane;
                if (COUNT && LM) {
                    const bool lm_copy = alive[s] && len[s] >= L + 1;  // decode.py:157
                    const bool lm_ext = alive[s] && len[s] >= L;       // decode.py:180
                    n_lookup += __popc(__ballot_sync(kFull, lm_copy)) + __popc(__ballot_sync(kFull, lm_ext));
                    n_combine += __popc(__ballot_sync(kFull, lm_copy && gcopy[s] && fgate)) +
                                 __popc(__ballot_sync(kFull, lm_ext && gext[s] && fgate));
                }
                // COPY (decode.py:150-175); the empty labeling and dead beams have pnb == 0
                dl[s] = rec[last[s]];
                if (LM && gcopy[s] && fgate) dl[s] = __dmul_rn(__dadd_rn(rcopy[s], rec[6 + last[s]]), rec[11]);
                npnb[s] = __dmul_rn(pnb[s], dl[s]);
                npb[s] = __dmul_rn(ptot[s], P4);
                nptot[s] = __dadd_rn(npb[s], npnb[s]);
                *reinterpret_cast<double2 *>(&sm.ex[b * 2]) = make_double2(ptot[s], pb[s]);
            }
            __syncwarp();
            // MERGE copy(X) with extend(parent(X), last(X)): the parent's extension by last(X) is
            // (pr_blank or pr_total of the parent) x this beam's own copy emission
            unsigned long long kcopy[BPL];
#pragma unroll
            for (int s = 0; s < BPL; ++s) {
                if (alive[s] && plane[s] >= 0) {
                    const double v = __dmul_rn(sm.ex[plane[s] * 2 + prep[s]], dl[s]);
                    npnb[s] = __dadd_rn(npnb[s], v);
                    nptot[s] = __dadd_rn(nptot[s], v);
                }
                kcopy[s] = (unsigned long long)__double_as_longlong(nptot[s]);
                sm.key[s * 32 + lane] = kcopy[s];  // zero for a dead beam
            }
            __syncwarp();

            // SELECT (decode.py:145, 35-39): does the order still hold, can any extension compete?
            bool ok = true;
#pragma unroll
            for (int s = 0; s < BPL; ++s)
                if (alive[s] && succ[s] != s * 32 + lane) {
                    // (bit-equal scores of two readings of one old ambiguity can last for the rest of the read:
                    // then the dict insertion positions decide, see decode.cu)
                    const unsigned long long ks = sm.key[succ[s]];
                    ok = ok && (kcopy[s] > ks || (kcopy[s] == ks && tie_ok[s]));
                }
            const unsigned long long kworst = sm.key[last_b];
            if (COUNT && fgate) n_diag += 1ull << 32;
            if (COUNT) n_diag += 1;  // a frame that has to look at extensions
            const bool order_ok = __all_sync(kFull, ok);

            RADIAN_LAP(3);
            // EXTEND (decode.py:177-201)
            unsigned long long ke[BPL][4];
            bool lost = false;  // see below
            {
                const double2 P01 = *reinterpret_cast<const double2 *>(rec);
                const double2 P23 = *reinterpret_cast<const double2 *>(rec + 2);
                if (LM && fgate) cp_async_wait_all();  // the rows of beams created in the frames before (uniform)
#pragma unroll
                for (int s = 0; s < BPL; ++s) {
                    const int b = s * 32 + lane;
                    double d0 = P01.x, d1 = P01.y, d2 = P23.x, d3 = P23.y;
                    if (LM && gext[s] && fgate) {
                        const double2 q01 = *reinterpret_cast<const double2 *>(rec + 6);
                        const double2 q23 = *reinterpret_cast<const double2 *>(rec + 8);
                        const double Sh = rec[11];
                        const double2 r01 = *reinterpret_cast<const double2 *>(&sm.row[b * 4]);
                        const double2 r23 = *reinterpret_cast<const double2 *>(&sm.row[b * 4 + 2]);
                        d0 = __dmul_rn(__dadd_rn(r01.x, q01.x), Sh);
                        d1 = __dmul_rn(__dadd_rn(r01.y, q01.y), Sh);
                        d2 = __dmul_rn(__dadd_rn(r23.x, q23.x), Sh);
                        d3 = __dmul_rn(__dadd_rn(r23.y, q23.y), Sh);
                    }
                    // a repeated symbol continues only paths that ended in a blank (decode.py:192-195);
                    // for the empty labeling (last = 0 by convention) pb == ptot
                    ke[s][0] = (unsigned long long)__double_as_longlong(__dmul_rn(last[s] == 0 ? pb[s] : ptot[s], d0));
                    ke[s][1] = (unsigned long long)__double_as_longlong(__dmul_rn(last[s] == 1 ? pb[s] : ptot[s], d1));
                    ke[s][2] = (unsigned long long)__double_as_longlong(__dmul_rn(last[s] == 2 ? pb[s] : ptot[s], d2));
                    ke[s][3] = (unsigned long long)__double_as_longlong(__dmul_rn(last[s] == 3 ? pb[s] : ptot[s], d3));
                    if (na < bw && alive[s]) {
                        // with room in the beam every extension is kept: one that underflowed would be
                        // ranked as a zero
                        const double dd[4] = {d0, d1, d2, d3};
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            lost = lost || (((km[s] >> (8 * c + 7)) & 1u) && (last[s] == c ? pb[s] : ptot[s]) != 0.0 &&
                                            dd[c] != 0.0 && (uint32_t)(ke[s][c] >> 32) < 0x00100000u);
                    }
                }
            }
            {
                // a kept beam (or, with room in the beam, an extension) that is not a normal number
                // although its factors are not zero: the spread of the beam exceeds what the rescaled
                // float64 scores can express
#pragma unroll
                for (int s = 0; s < BPL; ++s) {
                    bool nz = (ptot[s] != 0.0 && P4 != 0.0) || (pnb[s] != 0.0 && dl[s] != 0.0);
                    if (alive[s] && plane[s] >= 0) nz = nz || (sm.ex[plane[s] * 2 + prep[s]] != 0.0 && dl[s] != 0.0);
                    lost = lost || (alive[s] && nz && (uint32_t)(kcopy[s] >> 32) < 0x00100000u);
                }
                if (__any_sync(kFull, lost)) {
                    status = RADIAN_READ_RANGE;
                    break;
                }
            }
            unsigned long long tau = 0ull;  // with room left in the beam every extension is a candidate
            if (na >= bw) {
                if (order_ok) {
                    tau = kworst;
                } else {
                    unsigned long long mn = ~0ull;
#pragma unroll
                    for (int s = 0; s < BPL; ++s)
                        if (alive[s] && kcopy[s] < mn) mn = kcopy[s];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const unsigned long long x = __shfl_xor_sync(kFull, mn, o);
                        mn = x < mn ? x : mn;
                    }
                    tau = mn;
                }
            }
            bool comp[BPL][4];
            bool anyc = false;
#pragma unroll
            for (int s = 0; s < BPL; ++s)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    comp[s][c] = ((km[s] >> (8 * c + 7)) & 1u) && ke[s][c] >= tau;
                    anyc = anyc || comp[s][c];
                }
            const bool need = !order_ok || __any_sync(kFull, anyc);
            need_ = need;
            RADIAN_LAP(4);

            if (!need) {
#pragma unroll
                for (int s = 0; s < BPL; ++s)
                    if (alive[s]) {
                        ptot[s] = nptot[s];
                        pnb[s] = npnb[s];
                        pb[s] = npb[s];
                    }
            } else {
                // ---- candidate list: copies at [0,NB) by beam id, competing extensions behind
#pragma unroll
                for (int s = 0; s < BPL; ++s) sm.srank[s * 32 + lane] = (uint16_t)rank[s];
                __syncwarp();
#pragma unroll
                for (int s = 0; s < BPL; ++s) {
                    int pos_copy = 5 * rank[s];
                    if (alive[s] && plane[s] >= 0) {
                        // a merged copy keeps the earlier of its two insertion positions
                        const int pp = 5 * (int)sm.srank[plane[s]] + 1 + last[s];
                        pos_copy = pp < pos_copy ? pp : pos_copy;
                    }
                    sm.pos[s * 32 + lane] = alive[s] ? (uint16_t)pos_copy : kPosInvalid;
                }
                // (the order of the list is free: the insertion positions are explicit.  Every lane lists
                // its own extensions one after the other, behind those of the lanes before it)
                int n_ext = 0;
                {
                    int mine = 0;
#pragma unroll
                    for (int s = 0; s < BPL; ++s)
#pragma unroll
                        for (int c = 0; c < 4; ++c) mine += comp[s][c] ? 1 : 0;
                    int incl = mine;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int y = __shfl_up_sync(kFull, incl, o);
                        if (lane >= o) incl += y;
                    }
                    n_ext = __shfl_sync(kFull, incl, 31);
                    int ci = NB + incl - mine;
#pragma unroll
                    for (int s = 0; s < BPL; ++s)
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            if (comp[s][c]) {
                                sm.key[ci] = ke[s][c];
                                sm.k32[ci] = (uint32_t)(ke[s][c] >> 32);
                                sm.pos[ci] = (uint16_t)(5 * rank[s] + 1 + c);
                                sm.src[ci] = (uint16_t)((s * 32 + lane) * 4 + c);
                                ++ci;
                            }
                }
                const int m = NB + n_ext;
                __syncwarp();
                RADIAN_LAP(5);
                // ---- incremental ranks: a copy moves down from its rank among the copies by the
                // number of extensions that outrank it, and an extension ranks behind the copies and
                // extensions above it.  Counted on the high words; any equal pair of high words
                // sends the frame to the exact ranking below.
                uint32_t kc32[BPL];
                int base[BPL];  // rank of a copy among the copies
#pragma unroll
                for (int s = 0; s < BPL; ++s) {
                    kc32[s] = alive[s] ? (uint32_t)(kcopy[s] >> 32) : 0u;
                    sm.k32[s * 32 + lane] = kc32[s];
                    base[s] = rank[s];
                }
                __syncwarp();
                bool copies_ok = order_ok;
                if (!order_ok) {
                    // the copies changed order: ranked among themselves on the high words, which is
                    // exact when they are all distinct; the sum of the ranks proves it (a tie makes it
                    // fall short of na(na-1)/2) and otherwise the exact ranking below takes over
                    const uint4 *kv = reinterpret_cast<const uint4 *>(sm.k32);
                    int ssum = 0;
#pragma unroll
                    for (int s = 0; s < BPL; ++s) {
                        int cnt = 0;
                        const uint32_t k = kc32[s];
#pragma unroll 4
                        for (int j4 = 0; j4 < NB / 4; ++j4) {
                            const uint4 q = kv[j4];
                            cnt += (q.x > k) + (q.y > k) + (q.z > k) + (q.w > k);
                        }
                        base[s] = cnt;
                        ssum += alive[s] ? cnt : 0;
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) ssum += __shfl_xor_sync(kFull, ssum, o);
                    copies_ok = ssum == na * (na - 1) / 2;
                }
                bool inc = copies_ok && n_ext <= 2 * NB;
                if (inc) {
                    // The copies' high words in rank order are a descending array: an extension finds the
                    // number of copies above it by bisection and counts the extensions above it four at
                    // a time; a copy of rank r among the copies moves down by the number of extensions
                    // that have at most r copies above them (a histogram and its prefix sums).
                    uint32_t *const skey = sm.sctx;   // (staging array of the creation step: free here)
                    uint32_t *const hist = sm.kill;   // (rebuilt after the creation step: free here)
#pragma unroll
                    for (int s = 0; s < BPL; ++s) {
                        if (alive[s]) skey[base[s]] = kc32[s];
                        hist[s * 32 + lane] = 0u;
                    }
                    if (lane < 4) sm.k32[m + lane] = 0u;  // the extensions' words padded to a multiple of four
                    __syncwarp();
                    const uint4 *ev = reinterpret_cast<const uint4 *>(&sm.k32[NB]);
                    const int n4 = (n_ext + 3) >> 2;
                    int rsum = 0;
                    bool tie = false;
                    for (int ci = NB + lane; ci < m; ci += 32) {
                        const uint32_t k = sm.k32[ci];
                        int c = 0;  // copies above: the largest c with skey[c - 1] > k
#pragma unroll
                        for (int step = NB; step > 0; step >>= 1) {
                            const int tt = c + step;
                            const uint32_t v = skey[(tt < NB ? tt : NB) - 1];
                            if (tt <= na && v > k) c = tt;
                        }
                        if (c < na) {
                            tie = tie || skey[c] == k;  // a copy with my high word: the exact ranking decides
                            atomicAdd(&hist[c], 1u);
                        }
                        int x = 0;
#pragma unroll 2  // (A/B on B200, profiles/r2_ab_wide2.log: +5.6 % at width 64, +2.6 % at 128)
                        for (int j = 0; j < n4; ++j) {
                            const uint4 q = ev[j];
                            x += (q.x > k) + (q.y > k) + (q.z > k) + (q.w > k);
                        }
                        sm.rnk[ci] = (uint16_t)(c + x);
                        rsum += c + x;
                    }
                    __syncwarp();
                    // inclusive prefix sums of the histogram, BPL consecutive entries per lane
                    {
                        uint32_t h[BPL];
                        uint32_t run = 0;
#pragma unroll
                        for (int j = 0; j < BPL; ++j) {
                            run += hist[lane * BPL + j];
                            h[j] = run;
                        }
                        uint32_t incl = run;
#pragma unroll
                        for (int o = 1; o < 32; o <<= 1) {
                            const uint32_t y = __shfl_up_sync(kFull, incl, o);
                            if (lane >= o) incl += y;
                        }
                        const uint32_t excl = incl - run;
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < BPL; ++j) hist[lane * BPL + j] = excl + h[j];
                    }
                    __syncwarp();
#pragma unroll
                    for (int s = 0; s < BPL; ++s) {
                        const int r = alive[s] ? base[s] + (int)hist[base[s]] : 0xffff;
                        sm.rnk[s * 32 + lane] = (uint16_t)r;
                        rsum += alive[s] ? r : 0;
                    }
                    // two extensions with one high word leave the sum of all ranks short of mv(mv-1)/2
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) rsum += __shfl_xor_sync(kFull, rsum, o);
                    const int mv = na + n_ext;
                    if (__any_sync(kFull, tie) || rsum != mv * (mv - 1) / 2) inc = false;
                    __syncwarp();
                }
                RADIAN_LAP(6);
                // ---- exact ranks: (float64 bits desc, dict insertion position asc)
                if (!inc && !COUNT) {
                    // counted on the high words, four per load; only a candidate that shares its high word
                    // with another one (or is zero, like the dead copies) goes through all 64 bits
                    const uint4 *kv = reinterpret_cast<const uint4 *>(sm.k32);
                    const int m4 = m >> 2;
                    for (int ci = lane; ci < m; ci += 32) {
                        const uint16_t p = sm.pos[ci];
                        int cnt = 0xffff;
                        if (p != kPosInvalid) {
                            const unsigned long long k = sm.key[ci];
                            const uint32_t kh = (uint32_t)(k >> 32);
                            int gt = 0, eq = 0;
                            for (int j = 0; j < m4; ++j) {
                                const uint4 q = kv[j];
                                gt += (q.x > kh) + (q.y > kh) + (q.z > kh) + (q.w > kh);
                                eq += (q.x == kh) + (q.y == kh) + (q.z == kh) + (q.w == kh);
                            }
                            for (int j = m4 * 4; j < m; ++j) {
                                const uint32_t q = sm.k32[j];
                                gt += q > kh;
                                eq += q == kh;
                            }
                            cnt = gt;
                            if (eq > 1) {
                                cnt = 0;
                                for (int j = 0; j < m; ++j) {
                                    const uint16_t pj = sm.pos[j];
                                    const unsigned long long kj = sm.key[j];
                                    cnt += (pj != kPosInvalid) && (kj > k || (kj == k && pj < p));
                                }
                            }
                        }
                        sm.rnk[ci] = (uint16_t)cnt;
                    }
                } else if (!inc) {
                    // (COUNT instantiation: every pair, which also finds the near ties it reports)
                    bool near = false;  // two candidates within 2^-40 of each other (see decode.cu)
                    for (int ci = lane; ci < m; ci += 32) {
                        const uint16_t p = sm.pos[ci];
                        int cnt = 0xffff;
                        if (p != kPosInvalid) {
                            const unsigned long long k = sm.key[ci];
                            cnt = 0;
                            for (int j = 0; j < m; ++j) {
                                const uint16_t pj = sm.pos[j];
                                const unsigned long long kj = sm.key[j];
                                cnt += (pj != kPosInvalid) && (kj > k || (kj == k && pj < p));
                                if (COUNT) near = near || (j != ci && pj != kPosInvalid && k != 0ull && kj - k + 4096ull < 8192ull);
                            }
                        }
                        sm.rnk[ci] = (uint16_t)cnt;
                    }
                    if (COUNT) n_tie += __any_sync(kFull, near) ? 1 : 0;
                }
                __syncwarp();
                RADIAN_LAP(7);
                bool survive[BPL];
                unsigned survb[BPL], evb[BPL];
                int new_rank[BPL];
                int n_surv = 0;
#pragma unroll
                for (int s = 0; s < BPL; ++s) {
                    new_rank[s] = (int)sm.rnk[s * 32 + lane];
                    survive[s] = alive[s] && new_rank[s] < bw;
                    survb[s] = __ballot_sync(kFull, survive[s]);
                    evb[s] = __ballot_sync(kFull, alive[s] && !survive[s]);
                    n_surv += __popc(survb[s]);
                }
                int n_new = 0;
                for (int base = NB; base < m; base += 32) {
                    const int ci = base + lane;
                    const bool isnew = ci < m && sm.rnk[ci] < bw;
                    const unsigned bal = __ballot_sync(kFull, isnew);
                    if (isnew) sm.newlist[n_new + __popc(bal & below)] = (uint16_t)ci;
                    n_new += __popc(bal);
                }

                RADIAN_LAP(8);
                if (n_new > 0) {
                    // ---- stage what the new beams inherit from their parents
#pragma unroll
                    for (int s = 0; s < BPL; ++s) {
                        const int b = s * 32 + lane;
                        sm.sctx[b] = ctx[s];
                        sm.slen[b] = len[s];
                        sm.snode()[b] = node[s];
                        sm.sh()[b] = len[s] > 0 ? hash_step(hp[s], last[s]) : kHashEmpty;
                        sm.sgext[b] = (uint8_t)gext[s];
                        sm.slast[b] = (uint8_t)last[s];
                    }
                    if (LM) cp_async_wait_all();  // every lane's rows have landed before a child reads them
                    __syncwarp();
                    bool take[BPL];
                    int ford[BPL], item[BPL];
                    double p_r[BPL];
                    int fbase = 0;
                    int keyctx = -1, keyrank = 0x7fff;  // best-ranked new beam of this lane whose context the model lacks
#pragma unroll
                    for (int s = 0; s < BPL; ++s) {
                        const unsigned freeb = ~survb[s];
                        ford[s] = fbase + __popc(freeb & below);
                        fbase += __popc(freeb);
                        take[s] = !survive[s] && ford[s] < n_new;
                        item[s] = take[s] ? (int)sm.newlist[ford[s]] : 0;
                        p_r[s] = 0.0;
                        if (LM && take[s]) {
                            const int sc = (int)sm.src[item[s]];
                            if (sm.sgext[sc >> 2]) p_r[s] = sm.row[sc];  // row[parent*4 + c]
                        }
                    }
                    __syncwarp();  // all reads of parent rows done before any row is replaced
#pragma unroll
                    for (int s = 0; s < BPL; ++s) {
                        const int b = s * 32 + lane;
                        if (survive[s]) {
                            ptot[s] = nptot[s];
                            pnb[s] = npnb[s];
                            pb[s] = npb[s];
                            rank[s] = new_rank[s];
                            if (plane[s] >= 0) {
                                unsigned ev = 0;
#pragma unroll
                                for (int s2 = 0; s2 < BPL; ++s2)
                                    if ((plane[s] >> 5) == s2) ev = evb[s2];
                                if ((ev >> (plane[s] & 31)) & 1u) plane[s] = -1;
                            }
                        } else if (take[s]) {
                            const int sc = (int)sm.src[item[s]];
                            const int pbm = sc >> 2;
                            const int c = sc & 3;
                            const double scv = __longlong_as_double((long long)sm.key[item[s]]);
                            ptot[s] = scv;
                            pnb[s] = scv;
                            pb[s] = 0.0;
                            rank[s] = (int)sm.rnk[item[s]];
                            node[s] = top + ford[s];
                            len[s] = sm.slen[pbm] + 1;
                            ctx[s] = (sm.sctx[pbm] << 2) | (uint32_t)c;
                            last[s] = c;
                            hp[s] = sm.sh()[pbm];
                            unsigned sv = 0;
#pragma unroll
                            for (int s2 = 0; s2 < BPL; ++s2)
                                if ((pbm >> 5) == s2) sv = survb[s2];
                            plane[s] = ((sv >> (pbm & 31)) & 1u) ? pbm : -1;
                            prep[s] = (c == (int)sm.slast[pbm]) ? 1 : 0;
                            alive[s] = true;
                            arena[node[s]] = ((uint32_t)sm.snode()[pbm] << 2) | (uint32_t)c;
                            if (LM) {
                                gcopy[s] = sm.sgext[pbm] != 0;
                                rcopy[s] = p_r[s];
                                gext[s] = false;
                                rprov[s] = false;
                                if (len[s] >= L) {
                                    rprov[s] = true;  // the row is in flight from here on
                                    const uint32_t ci = ctx[s] & ctx_mask;
                                    const uint32_t gwd = __ldg(a.gate + (ci >> 5));
                                    if (a.miss != nullptr && ((__ldg(a.miss + (ci >> 5)) >> (ci & 31u)) & 1u) &&
                                        rank[s] < keyrank) {
                                        keyrank = rank[s];
                                        keyctx = (int)ci;
                                    }
                                    const double *row = a.table + (size_t)ci * 4;
                                    cp_async<16>(&sm.row[b * 4], row);
                                    cp_async<16>(&sm.row[b * 4 + 2], row + 2);
                                    gext[s] = (gwd >> (ci & 31u)) & 1u;
                                }
                            }
                        } else {
                            alive[s] = false;
                            ptot[s] = pnb[s] = pb[s] = 0.0;
                        }
                    }
                    RADIAN_LAP(9);
                    top += n_new;
                    na = n_surv + n_new;
                    if (LM && a.miss != nullptr && t + 1 < T) {
                        // a kept beam whose extend-context the model does not hold: KeyError at
                        // lm[context] when the reference processes it in the next frame, best rank
                        // first (decode.py:83); after the last frame it never looks
                        int rmin = keyrank;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) {
                            const int x = __shfl_xor_sync(kFull, rmin, o);
                            rmin = x < rmin ? x : rmin;
                        }
                        if (rmin != 0x7fff) {
                            if (keyrank == rmin) a.out_len[read] = keyctx;  // the context index, for the message
                            status = RADIAN_READ_KEY_ERROR;
                        }
                    }
                    __syncwarp();  // parents' staged values have been consumed
                    // a surviving beam whose parent labeling was just (re)created points at it again
#pragma unroll
                    for (int s = 0; s < BPL; ++s)
                        if (take[s]) {
                            const int b = s * 32 + lane;
                            sm.sh()[b] = hash_step(hp[s], last[s]);
                            sm.slen[b] = len[s];
                            sm.slast[b] = (uint8_t)last[s];
                        }
                    __syncwarp();
                    RADIAN_LAP(10);
                    {
                        // an open-addressing table of the new beams by labeling hash (4 NB slots in the
                        // high-word array, whose ranking job is done); every orphan looks its parent up
                        constexpr unsigned HM = 4 * NB - 1;
                        uint32_t *const tab = sm.k32;
#pragma unroll
                        for (int j = 0; j < BPL; ++j)
                            *reinterpret_cast<uint4 *>(&tab[(j * 32 + lane) * 4]) = make_uint4(~0u, ~0u, ~0u, ~0u);
                        __syncwarp();
#pragma unroll
                        for (int s = 0; s < BPL; ++s)
                            if (take[s]) {
                                unsigned i = (unsigned)(hash_step(hp[s], last[s]) >> 20) & HM;
                                while (atomicCAS(&tab[i], ~0u, (uint32_t)(s * 32 + lane)) != ~0u) i = (i + 1) & HM;
                            }
                        __syncwarp();
#pragma unroll
                        for (int s = 0; s < BPL; ++s)
                            if (survive[s] && plane[s] < 0 && len[s] > 0) {
                                unsigned i = (unsigned)(hp[s] >> 20) & HM;
                                while (true) {
                                    const uint32_t zb = *(volatile uint32_t *)&tab[i];
                                    if (zb == ~0u) break;
                                    if (sm.sh()[zb] == hp[s] && sm.slen[zb] + 1 == len[s]) {
                                        plane[s] = (int)zb;
                                        prep[s] = ((int)sm.slast[zb] == last[s]) ? 1 : 0;
                                        break;
                                    }
                                    i = (i + 1) & HM;
                                }
                            }
                    }
                    RADIAN_LAP(11);
                    // the beam set changed: refresh which extensions are merged into a live child
#pragma unroll
                    for (int s = 0; s < BPL; ++s) sm.kill[s * 32 + lane] = 0u;
                    __syncwarp();
#pragma unroll
                    for (int s = 0; s < BPL; ++s)
                        if (alive[s] && plane[s] >= 0)
                            reinterpret_cast<uint8_t *>(sm.kill)[plane[s] * 4 + last[s]] = 0x80;
                    __syncwarp();
#pragma unroll
                    for (int s = 0; s < BPL; ++s) {
                        km[s] = alive[s] ? (0x80808080u & ~sm.kill[s * 32 + lane]) : 0u;
                        rmax[s] = kNoRmaxW;  // the merge mask (or the beam) changed
                    }
                } else {
#pragma unroll
                    for (int s = 0; s < BPL; ++s)
                        if (alive[s]) {  // nothing new entered: every beam stays
                            ptot[s] = nptot[s];
                            pnb[s] = npnb[s];
                            pb[s] = npb[s];
                            rank[s] = new_rank[s];
                        }
                }
                RADIAN_LAP(12);
                // successor of every beam, best and worst beam
                __syncwarp();
#pragma unroll
                for (int s = 0; s < BPL; ++s)
                    if (alive[s]) sm.byrank[rank[s]] = (uint16_t)(s * 32 + lane);
                __syncwarp();
#pragma unroll
                for (int s = 0; s < BPL; ++s)
                    succ[s] = (alive[s] && rank[s] + 1 < na) ? (int)sm.byrank[rank[s] + 1] : s * 32 + lane;
                first = (int)sm.byrank[0];
                last_b = (int)sm.byrank[na - 1];
                // dict insertion positions the next frame's copies will have, from the new ranks: is a
                // successor with exactly my score rightly behind me?
#pragma unroll
                for (int s = 0; s < BPL; ++s) sm.srank[s * 32 + lane] = (uint16_t)rank[s];
                __syncwarp();
#pragma unroll
                for (int s = 0; s < BPL; ++s) {
                    int npos = 5 * rank[s];
                    if (alive[s] && plane[s] >= 0) {
                        const int pp = 5 * (int)sm.srank[plane[s]] + 1 + last[s];
                        npos = pp < npos ? pp : npos;
                    }
                    sm.pos[s * 32 + lane] = (uint16_t)npos;
                }
                __syncwarp();
#pragma unroll
                for (int s = 0; s < BPL; ++s) tie_ok[s] = alive[s] && sm.pos[s * 32 + lane] < sm.pos[succ[s]];
                __syncwarp();
            }
            RADIAN_LAP(14);
            } while (0);
            ++t;
            refresh();
            RADIAN_LAP(15);
#ifdef RADIAN_WIDE_PROBE
            if (lane == 0) sm.lap[16 + (need_ ? 1 : 0)] += 1;
#endif
        }
        cp_async_wait_all();
        __syncwarp();

#ifdef RADIAN_WIDE_PROBE
        __syncwarp();
        if (lane < 24) atomicAdd(&g_lap[lane], sm.lap[lane]);
#endif
        // ------------------------------------------------------------ end of read
        if (status == 0) {
            const long long seq_off = a.seq_offsets[read];
            const long long seq_cap = a.seq_offsets[read + 1] - seq_off;
            int second = -1;
#pragma unroll
            for (int s = 0; s < BPL; ++s) {
                const int x = __shfl_sync(kFull, succ[s], first & 31);
                if ((first >> 5) == s) second = x;
            }
            {
                // the best two beams, exactly: if they share their high word, their low words may have
                // crossed since the order was last checked on all 64 bits (q_inc = 0, see REFRESH)
                bool t1 = false;
#pragma unroll
                for (int s = 0; s < BPL; ++s) {
                    sm.key[s * 32 + lane] = (unsigned long long)__double_as_longlong(ptot[s]);
                    t1 = t1 || (s * 32 + lane == first && tie_ok[s]);
                }
                t1 = __any_sync(kFull, t1);
                __syncwarp();
                const unsigned long long p1 = sm.key[first], p2 = sm.key[second];
                if (second != first && (p2 > p1 || (p2 == p1 && !t1))) {
                    const int x = first;
                    first = second;
                    second = x;
                }
            }
#pragma unroll
            for (int s = 0; s < BPL; ++s) {
                const int b = s * 32 + lane;
                if (b == first) {
                    const long long n = len[s];
                    int st = 0;
                    if (n > seq_cap) st = RADIAN_READ_SEQ_OVERFLOW;
                    int c = node[s];
                    for (long long i = n - 1; i >= 0; --i) {
                        const uint32_t w = arena[c] & 0x7fffffffu;
                        if (i < seq_cap) a.out_seq[seq_off + i] = (uint8_t)(w & 3u);
                        c = (int)(w >> 2);
                    }
                    a.out_len[read] = n;
                    a.out_score[2 * read] = final_log_score(ptot[s], kacc);
                    if (second == first) a.out_score[2 * read + 1] = NAN;
                    a.out_status[read] = st;
                    if (a.out_counters) {
                        a.out_counters[4 * read] = n_lookup;
                        a.out_counters[4 * read + 1] = n_combine;
                        a.out_counters[4 * read + 2] = n_tie;
                        a.out_counters[4 * read + 3] = n_diag;
                    }
                }
                if (second != first && b == second)
                    a.out_score[2 * read + 1] = final_log_score(ptot[s], kacc);
            }
        } else if (lane == 0) {
            if (status != RADIAN_READ_KEY_ERROR) a.out_len[read] = 0;  // (KeyError: holds the context index)
            a.out_score[2 * read] = NAN;
            a.out_score[2 * read + 1] = NAN;
            a.out_status[read] = status;
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------ host side

template <int BPL, bool LM, typename PT>
static const void *wide_ptr(bool count)
{
    return count ? (const void *)decode_wide_kernel<BPL, LM, PT, true>
                 : (const void *)decode_wide_kernel<BPL, LM, PT, false>;
}

int wide_pick(int beam_width, bool lm, bool f64, bool count, const void **kernel, size_t *smem_bytes)
{
    const int bpl = beam_width <= 64 ? 2 : 4;
#define RADIAN_WIDE(B)                                                                           \
    if (bpl == B) {                                                                              \
        if (lm) {                                                                                \
            *kernel = f64 ? wide_ptr<B, true, double>(count) : wide_ptr<B, true, float>(count);   \
            *smem_bytes = kWideWarps * (f64 ? sizeof(WideSmem<B, true, double>) : sizeof(WideSmem<B, true, float>)); \
        } else {                                                                                 \
            *kernel = f64 ? wide_ptr<B, false, double>(count) : wide_ptr<B, false, float>(count); \
            *smem_bytes = kWideWarps * (f64 ? sizeof(WideSmem<B, false, double>) : sizeof(WideSmem<B, false, float>)); \
        }                                                                                        \
        return kWideWarps;                                                                       \
    }
    RADIAN_WIDE(2)
    RADIAN_WIDE(4)
#undef RADIAN_WIDE
    return 0;
}

#ifdef RADIAN_WIDE_PROBE
}  // namespace radian
extern "C" int radian_debug_wide_laps(unsigned long long *out, int reset)
{
    cudaError_t e = cudaMemcpyFromSymbol(out, radian::g_lap, sizeof(unsigned long long) * 24);
    if (e == cudaSuccess && reset) {
        unsigned long long z[24] = {0};
        e = cudaMemcpyToSymbol(radian::g_lap, z, sizeof(z));
    }
    return (int)e;
}
namespace radian {
#endif
}  // namespace radian
