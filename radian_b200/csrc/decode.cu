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
#include "decode_common.cuh"

namespace radian {

constexpr int kWarpsPerBlock = 4;
constexpr int kNoRmax = (int)0x80000000;
// resident CTAs per SM asked from ptxas (A/B on B200, profiles/r1_minblocks_ab.txt): 6 CTAs =
// 24 warps at <= 80 registers once the tile prefetch and the RNA rows moved to shared memory
// (A/B on B200, scripts/ab_variants.sh: trading the pb/ptot selects of the extension scores for one
// more float64 multiply lost 6 %: the FP64 pipe is the scarce unit, so the frame loop avoids it)
#ifndef RADIAN_MIN_BLOCKS
#define RADIAN_MIN_BLOCKS 6
#endif

template <int G, bool LM, typename PT>
struct __align__(16) GroupSmem {
    // doubles per frame (decode_common.cuh, EXT layout): P0..P4, gate, q0..q3, S, S/2, then the
    // integer words of the quiet-frame test; 144 / 80 bytes per frame keep the record stores
    // of neighbouring lanes on different banks
    static constexpr int REC = LM ? 18 : 10;
    static constexpr int NK = (2 * G > 32) ? 2 * G : 32;  // candidate slots of the fast ranking path
    double rec[G * REC];
    double row[LM ? G * 4 : 4];       // RNA table row of every lane's extend-context (cp.async target)
    PT raw[G * 5];                    // next tile of posterior rows, landed by cp.async
    double ex[G * 2];                 // {pr_total, pr_blank} of every lane before the frame (copy/extend merge)
    unsigned long long key[5 * G];    // candidate list: [0,G) copies by lane, [G,..) extensions
    uint32_t k32[NK];                 // high words of the candidate scores (0 = empty slot)
    uint32_t kill[G];                 // byte c of word l: extension (l,c) merged into a copy
    uint16_t pos[5 * G];              // dict insertion position of the candidate
    uint8_t src[5 * G];               // lane*4+c of an extension candidate
    uint8_t rnk[5 * G];               // rank of the candidate
    uint8_t newlist[G];               // candidate indices of the new beams, in list order
    uint8_t lanerank[G];              // previous rank of every lane
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
    constexpr int NK = GroupSmem<G, LM, PT>::NK;
    constexpr int EPL = (NK - G) / G;  // extension slots ranked by each lane on the fast path
    constexpr unsigned GBITS = (G == 32) ? kFull : ((1u << G) - 1u);
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
    unsigned long long h = 0, hp = 0;
    uint32_t ctx = 0;
    int len = 0, node = 0, rank = 0, plane = -1, last = 0;
    int prep = 0;        // 1 if the live parent (plane) ends in the same symbol as this beam
    int rmax = kNoRmax;  // max high word of the unmerged entries of this beam's table row, or kNoRmax
    bool alive = false;
    double rcopy = 0;  // table value of this beam's last symbol in its copy-context; the row of the
                       // extend-context lives in sm.row[li*4..]
    bool gext = false, gcopy = false;
    int succ = 0;        // absolute lane of the beam ranked right after this one (own lane: none)
    uint32_t km = 0;     // byte c = 0x80: this lane holds a beam and its extension by c is a candidate
                         // of its own (not merged into a live child's copy); 0 for a dead lane
    // ---- per-read (group-uniform) state
    int read = -1, top = 0, old_top = 0, na = 0, status = 0, first_lane = 0, last_lane = 0;
    int T = 0, t = 0, pend = -1;  // pend: queue ticket of a read that has not landed yet
    long long kacc = 0;
    const PT *rp = (const PT *)a.post;
    unsigned long long n_lookup = 0, n_combine = 0, n_tie = 0;
    bool active = true;

    while (true) {
        // ------------------------------------------------------------ fetch a read
        {
            const bool want = active && read < 0;
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
                    const long long foff = a.frame_offsets[read];
                    T = (int)(a.frame_offsets[read + 1] - foff);
                    rp = (const PT *)a.post + foff * 5;
                    t = 0;
                    // initial beam: the empty labeling, pr_blank = pr_total = log 1 (decode.py:128-132)
                    alive = (li == 0);
                    ptot = alive ? 1.0 : 0.0;
                    pb = ptot;
                    pnb = 0.0;
                    h = 0x243F6A8885A308D3ull;
                    hp = 0;
                    ctx = 0;
                    len = 0;
                    node = 0;
                    rank = 0;
                    plane = -1;
                    prep = 0;
                    rmax = kNoRmax;
                    last = 0;
                    gext = gcopy = false;
                    succ = lane;
                    km = alive ? 0x80808080u : 0u;
                    first_lane = gshift;
                    last_lane = gshift;
                    top = 1;  // node 0 = the empty labeling
                    old_top = 1;
                    na = 1;
                    status = 0;
                    kacc = 0;
                    n_lookup = n_combine = n_tie = 0;
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

        for (int it = 0; it < nrun; ++it) {
            bool run = live && status == 0;  // group-uniform
            // -------------------------------------------------------- tile refill
            if ((it % G) == 0) {
                cp_async_wait_all();
                __syncwarp();
                if (run && tb + it + li < T) make_record<LM, true>(&sm.raw[li * 5], a.s_thr, &sm.rec[li * REC]);
                __syncwarp();
                if (run && tb + it + G + li < T) prefetch_row(&sm.raw[li * 5], rp, tb + it + G + li);
            }

            // -------------------------------------------------------- nursery collection
            // checked once per tile: a frame adds at most G nodes per read, a tile at most G*G
            if ((it % G) == 0 && __any_sync(kFull, run && (top + G * G > old_top + kNursery || top + G * G > cap))) {
                // every running group of the warp collects (early collection is harmless)
                // 1. mark nursery nodes reachable from a live beam (stop at the old generation or
                //    at a node somebody marked in an earlier step)
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
                    const unsigned bal = GBALLOT(mk);
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
                        arena[ni] = ((uint32_t)npar << 2) | (w & 3u);
                        fwd[i - old_top] = (uint32_t)ni;
                    }
                    cnt += __popc(bal);
                    __syncwarp();
                }
                if (run && alive && node >= old_top) node = (int)fwd[node - old_top];
                __syncwarp();
                if (run) {
                    old_top = cnt;
                    top = cnt;
                    if (top + G * G > cap) {
                        status = RADIAN_READ_TRIE_OVERFLOW;  // reported; remaining frames are skipped
                        run = false;
                        alive = false;
                        km = 0u;
                        succ = lane;
                        ptot = pnb = pb = 0.0;
                    }
                }
            }
            const bool av = alive && run;  // this lane holds a beam that takes part in this frame

            // RESCALE by an exact power of two when the best beam (the largest value of the group)
            // has fallen below 2^-256.  Checked every frame: a float32-derived probability is
            // >= 2^-149, so nothing gets near the float64 limit in between.  Idle groups hold zeros.
            {
                const int exb = __shfl_sync(kFull, __double2hiint(ptot), first_lane) >> 20;
                if (exb != 0 && exb < 1023 - 256) {
                    const double sc = __hiloint2double((2046 - exb) << 20, 0);
                    ptot *= sc;
                    pnb *= sc;
                    pb *= sc;
                    kacc += exb - 1023;
                }
            }

            // -------------------------------------------------------- one frame
            const double *rec = &sm.rec[(it % G) * REC];
            const int *reci = reinterpret_cast<const int *>(rec);
            const double P4 = rec[4];
            bool fgate = false;
            int hS = 0;
            if (LM) {
                const int2 gs = *reinterpret_cast<const int2 *>(reci + 32);
                fgate = gs.x != 0;
                hS = gs.y;
            }
            // (a dead lane computes on stale flags; all its scores are zero and stay zero)
            if (COUNT && LM) {
                const bool lm_copy = av && len >= L + 1;  // decode.py:157
                const bool lm_ext = av && len >= L;       // decode.py:180
                n_lookup += __popc(GBALLOT(lm_copy)) + __popc(GBALLOT(lm_ext));
                n_combine += __popc(GBALLOT(lm_copy && gcopy && fgate)) + __popc(GBALLOT(lm_ext && gext && fgate));
            }

            // COPY (decode.py:150-175)
            // the empty labeling and dead lanes have pnb == 0, so their copy needs no special case
            double dl_ = rec[last];
            // (gcopy implies len >= L+1 and gext implies len >= L: both are set when the beam is created)
            if (LM && gcopy && fgate) dl_ = __dmul_rn(__dadd_rn(rcopy, rec[6 + last]), rec[11]);  // decode.py:58-61
            double npnb = __dmul_rn(pnb, dl_);
            const double npb = __dmul_rn(ptot, P4);
            double nptot = __dadd_rn(npb, npnb);

            // MERGE copy(X) with extend(parent(X), last(X)): same dict key in the reference.  The
            // parent's extension by last(X) is (pr_blank or pr_total of the parent) x the emission
            // of last(X) in the parent's extend-context, and that emission is this beam's own
            // copy emission dl_ (same context, same gate, same table value), so only the parent's
            // two scores travel through shared memory.  Which pairs merge only changes when the
            // beam set changes: the pairing (plane, prep, km) is state.
            asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(ex_addr), "d"(ptot), "d"(pb) : "memory");
            __syncwarp();
            if (av && plane >= 0) {
                const double v = __dmul_rn(sm.ex[plane * 2 + prep], dl_);
                npnb = __dadd_rn(npnb, v);
                nptot = __dadd_rn(nptot, v);
            }

            // SELECT the best beam_width candidates (decode.py:145, 35-39).
            // Most frames change nothing but the scores: the rank order is kept as state (succ =
            // lane of the next-ranked beam) and re-validated with one compare per beam; an
            // extension matters only if it is not below the worst copy of a full beam.
            const unsigned long long kcopy = (unsigned long long)__double_as_longlong(nptot);
            const uint32_t kc32 = (uint32_t)(kcopy >> 32);  // zero on a dead lane: its scores are zero
            const uint32_t ksucc = __shfl_sync(kFull, kc32, succ);
            const uint32_t kworst = __shfl_sync(kFull, kc32, last_lane);
            const bool prune = (na >= bw);
            bool ranks_changed = false;
            {
                // QUIET frame (the common case, one vote): in every group of the warp the order of
                // the copies holds strictly, the beam is full and no extension, merged ones
                // excepted, can reach the worst copy.  The extensions are not computed for this:
                // with h(x) = high word of the float64 x, (h(x) >> 20) - 1023 + mantissa fraction
                // is a lower bound of log2 x that is short by at most 0.0861, so
                //   h(p) + h(d) - bias + slack < h(worst)  implies  p*d < worst
                // for slack >= 2 * 0.0861 * 2^20.  The emission d of an extension is P_c, or with
                // the model ((r_c + q_c)/2) * S <= max(r_c, q_c) * S (one more 0.0861).
                // (A/B on B200: 5 or 6 resident CTAs per SM make no difference any more, 9.09e9 vs
                // 9.06e9 frames/s; 4 lose 10 %.)
                // With the model the table part of the bound, max over the unmerged symbols of
                // h(r_c), is a per-beam constant (rmax): it is rebuilt lazily from the row in shared
                // memory after the beam was created or its merge mask changed.
                const bool gated = LM && gext && fgate;
                if (LM && gated && rmax == kNoRmax) {
                    cp_async_wait_all();  // the row gathered when this beam was created
                    const int4 ra = *reinterpret_cast<const int4 *>(&sm.row[li * 4]);      // r0 lo,hi r1 lo,hi
                    const int4 rb = *reinterpret_cast<const int4 *>(&sm.row[li * 4 + 2]);  // r2, r3
                    rmax = max(max(ra.y & (int)byte_sign_mask<0>(km), ra.w & (int)byte_sign_mask<1>(km)),
                               max(rb.y & (int)byte_sign_mask<2>(km), rb.w & (int)byte_sign_mask<3>(km)));
                }
                // high words of q0..q3 (gated lanes) or of P0..P3
                // (both loaded, then selected: one load whose address waits for `gated` is 1 % slower.
                // Other A/Bs on B200: pinning `lane` in a register to spare the per-frame re-read of
                // SR_TID costs more in register pressure than it saves, -2 %; refreshing the rescale
                // exponent where the scores are committed instead of at the frame start is a wash.)
                int4 hx = *reinterpret_cast<const int4 *>(reci + (LM ? 24 : 12));
                if (LM) {
                    const int4 hq = *reinterpret_cast<const int4 *>(reci + 28);
                    if (gated) hx = hq;
                }
                // slack: 2 * 0.0861 * 2^20 for p * P_c, one more 0.0861 for the max(r, q) * S bound
                const int kSlack = (LM && gated) ? 272000 : 181000;
                const int z0 = hx.x & (int)byte_sign_mask<0>(km);
                const int z1 = hx.y & (int)byte_sign_mask<1>(km);
                const int z2 = hx.z & (int)byte_sign_mask<2>(km);
                const int z3 = hx.w & (int)byte_sign_mask<3>(km);
                int zmax = max(max(z0, z1), max(z2, z3));
                if (LM && gated) zmax = max(zmax, rmax) + hS;
                const int ub = __double2hiint(ptot) + zmax + (kSlack - 0x3ff00000);
                const bool quiet = !run || ((kc32 > ksucc || succ == lane) && prune && kworst >= 0x00100000u &&
                                            ub < (int)kworst);
                if (__all_sync(kFull, quiet)) {
                    ptot = nptot;  // (a dead lane's new values are zero as well)
                    pnb = npnb;
                    pb = npb;
                    continue;
                }
            }

            // EXTEND (decode.py:177-201), only when some extension may matter
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
            const bool order_ok = GBALLOT(kc32 > ksucc || succ == lane) == GBITS;
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
            bool full;

            if (__any_sync(kFull, !order_ok)) {
                // some group's copies changed order: its worst copy is the minimum over the lanes
                uint32_t tmin = av ? kc32 : 0xffffffffu;
#pragma unroll
                for (int o = G / 2; o > 0; o >>= 1) {
                    const uint32_t x = __shfl_xor_sync(kFull, tmin, o);
                    tmin = x < tmin ? x : tmin;
                }
                if (!order_ok) tau = (prune && tmin > 1u) ? tmin : 1u;
                full = GBALLOT(run && xmax >= tau) != 0u;  // my group needs a full ranking
                // groups without a competing extension rank their copies on the high words alone
                sm.k32[li] = kc32;
                __syncwarp();
                int cc = 0;
                const uint4 *kv = reinterpret_cast<const uint4 *>(sm.k32);
#pragma unroll
                for (int j = 0; j < G / 4; ++j) {
                    const uint4 k4 = kv[j];
                    cc += (k4.x > kc32) + (k4.y > kc32) + (k4.z > kc32) + (k4.w > kc32);
                }
                int ssum = av ? cc : 0;
#pragma unroll
                for (int o = G / 2; o > 0; o >>= 1) ssum += __shfl_xor_sync(kFull, ssum, o);
                if (!full && !order_ok) {
                    if (ssum == na * (na - 1) / 2) {
                        if (av) rank = cc;
                        ranks_changed = true;
                    } else {
                        full = true;  // two copies share a high word: exact ranking below
                    }
                }
                __syncwarp();
            } else {
                full = GBALLOT(run && xmax >= tau) != 0u;
            }

            if (!__any_sync(kFull, full)) {
                // (a dead lane's new values are zero as well)
                ptot = nptot;
                pnb = npnb;
                pb = npb;
            } else {
                // all groups of the warp go through the ranking; one that did not ask for it has no
                // extension candidates and gets its current order back
                ranks_changed = ranks_changed || run;
#pragma unroll
                for (int e = 0; e < EPL; ++e) sm.k32[G + e * G + li] = 0u;
                __syncwarp();
                int n_ext = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const double ec = c == 0 ? e0 : c == 1 ? e1 : c == 2 ? e2 : e3;
                    const bool comp = full && (c == 0 ? x0 : c == 1 ? x1 : c == 2 ? x2 : x3) >= tau;
                    const unsigned bal = GBALLOT(comp);
                    if (comp) {
                        const int idx = G + n_ext + __popc(bal & belowg);
                        const unsigned long long kc = (unsigned long long)__double_as_longlong(ec);
                        sm.key[idx] = kc;
                        sm.pos[idx] = (uint16_t)(5 * rank + 1 + c);
                        sm.src[idx] = (uint8_t)(li * 4 + c);
                        if (idx < NK) sm.k32[idx] = (uint32_t)(kc >> 32);
                    }
                    n_ext += __popc(bal);
                }
                sm.k32[li] = kc32;
                sm.lanerank[li] = (uint8_t)rank;
                __syncwarp();
                const int m = G + n_ext;
                int new_rank = 255;
                bool fast = (m <= NK);
                {
                    // rank = number of candidates with a strictly larger high word.  Exact whenever
                    // the high words of the ranked candidates are all distinct, which the rank sum
                    // proves (any tie makes the sum fall short of mv(mv-1)/2).
                    uint32_t ke[EPL];
                    int ce[EPL];
#pragma unroll
                    for (int e = 0; e < EPL; ++e) {
                        ke[e] = sm.k32[G + e * G + li];
                        ce[e] = 0;
                    }
                    int cc = 0;
                    const uint4 *kv = reinterpret_cast<const uint4 *>(sm.k32);
                    // the copies, then only as many extension slots as some group of the warp
                    // filled (the others hold zero and count for nothing)
                    int jend = n_ext < NK - G ? n_ext : NK - G;
#pragma unroll
                    for (int o = 16; o >= G; o >>= 1) {
                        const int x = __shfl_xor_sync(kFull, jend, o);
                        jend = x > jend ? x : jend;
                    }
                    jend = (G + jend + 3) / 4;
#pragma unroll
                    for (int j = 0; j < G / 4; ++j) {
                        const uint4 k4 = kv[j];
                        cc += (k4.x > kc32) + (k4.y > kc32) + (k4.z > kc32) + (k4.w > kc32);
#pragma unroll
                        for (int e = 0; e < EPL; ++e)
                            ce[e] += (k4.x > ke[e]) + (k4.y > ke[e]) + (k4.z > ke[e]) + (k4.w > ke[e]);
                    }
                    for (int j = G / 4; j < jend; ++j) {
                        const uint4 k4 = kv[j];
                        cc += (k4.x > kc32) + (k4.y > kc32) + (k4.z > kc32) + (k4.w > kc32);
#pragma unroll
                        for (int e = 0; e < EPL; ++e)
                            ce[e] += (k4.x > ke[e]) + (k4.y > ke[e]) + (k4.z > ke[e]) + (k4.w > ke[e]);
                    }
                    int ssum = av ? cc : 0;
#pragma unroll
                    for (int e = 0; e < EPL; ++e)
                        if (e * G + li < n_ext) ssum += ce[e];
#pragma unroll
                    for (int o = G / 2; o > 0; o >>= 1) ssum += __shfl_xor_sync(kFull, ssum, o);
                    const int mv = na + n_ext;
                    fast = fast && (ssum == mv * (mv - 1) / 2);
                    if (fast) {
                        new_rank = av ? cc : 255;
#pragma unroll
                        for (int e = 0; e < EPL; ++e)
                            if (e * G + li < n_ext) sm.rnk[G + e * G + li] = (uint8_t)ce[e];
                    }
                }
                if (__any_sync(kFull, run && !fast)) {
                    // exact path: (score desc, dict insertion position asc) on the full float64 bits;
                    // a merged copy keeps the earlier of its two insertion positions
                    if (run && !fast) {
                        int pos_copy = 5 * rank;
                        if (av && plane >= 0) {
                            const int pp = 5 * (int)sm.lanerank[plane] + 1 + last;
                            pos_copy = pp < pos_copy ? pp : pos_copy;
                        }
                        sm.key[li] = kcopy;
                        sm.pos[li] = av ? (uint16_t)pos_copy : kPosInvalid;
                    }
                    __syncwarp();
                    bool near = false;
                    if (run && !fast) {
                        for (int idx = li; idx < m; idx += G) {
                            const uint16_t p = sm.pos[idx];
                            if (p != kPosInvalid) {
                                const unsigned long long k = sm.key[idx];
                                int cnt = 0;
                                for (int j = 0; j < m; ++j) {
                                    const uint16_t pj = sm.pos[j];
                                    const unsigned long long kj = sm.key[j];
                                    cnt += (pj != kPosInvalid) && (kj > k || (kj == k && pj < p));
                                    // two candidates within 2^-40 of each other: a decision that the
                                    // log-domain reference takes on its own rounding noise
                                    if (COUNT) near = near || (j != idx && pj != kPosInvalid && k != 0ull && kj - k + 4096ull < 8192ull);
                                }
                                sm.rnk[idx] = (uint8_t)(cnt > 255 ? 255 : cnt);
                            }
                        }
                    }
                    __syncwarp();
                    if (run && !fast) new_rank = av ? (int)sm.rnk[li] : 255;
                    if (COUNT) n_tie += (GBALLOT(near) != 0u);
                } else {
                    __syncwarp();
                }

                const bool survive = av && new_rank < bw;
                const unsigned evb = GBALLOT(av && !survive);
                const unsigned survb = GBALLOT(survive);
                const unsigned freeb = GBITS & ~survb;
                int mmax = m;
#pragma unroll
                for (int o = 16; o >= G; o >>= 1) {
                    const int x = __shfl_xor_sync(kFull, mmax, o);
                    mmax = x > mmax ? x : mmax;
                }
                int n_new = 0;
                for (int base = G; base < mmax; base += G) {
                    const int idx = base + li;
                    const bool isnew = run && idx < m && sm.rnk[idx] < bw;
                    const unsigned bal = GBALLOT(isnew);
                    if (isnew) sm.newlist[n_new + __popc(bal & belowg)] = (uint8_t)idx;
                    n_new += __popc(bal);
                }

                if (__any_sync(kFull, n_new > 0)) {
                    __syncwarp();
                    const int ford = __popc(freeb & belowg);
                    const bool take = run && !survive && ford < n_new;
                    const int item = take ? (int)sm.newlist[ford] : 0;
                    const int s = take ? (int)sm.src[item] : li * 4;
                    const int ls = (s >> 2) + gshift;
                    const int c = s & 3;
                    // parent state, read before anybody overwrites it
                    const uint32_t p_ctx = __shfl_sync(kFull, ctx, ls);
                    const int p_len = __shfl_sync(kFull, len, ls);
                    const int p_node = __shfl_sync(kFull, node, ls);
                    const unsigned long long p_h = __shfl_sync(kFull, h, ls);
                    const int p_last = __shfl_sync(kFull, last, ls);
                    double p_r = 0.0;
                    bool p_g = false;
                    if (LM) {
                        p_g = __shfl_sync(kFull, (int)gext, ls) != 0;
                        // every lane waits for its own row first; the parent's row is then complete
                        cp_async_wait_all();
                        __syncwarp();
                        if (take && p_g) p_r = sm.row[(ls - gshift) * 4 + c];
                        __syncwarp();  // all reads of parent rows done before any row is replaced
                    }
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
                        node = top + ford;
                        len = p_len + 1;
                        ctx = (p_ctx << 2) | (uint32_t)c;
                        last = c;
                        hp = p_h;
                        h = hash_step(p_h, c);
                        plane = ((survb >> (ls - gshift)) & 1u) ? (ls - gshift) : -1;
                        prep = (c == p_last) ? 1 : 0;
                        alive = true;
                        arena[node] = ((uint32_t)p_node << 2) | (uint32_t)c;
                        if (LM) {
                            gcopy = p_g;
                            rcopy = p_r;
                            gext = false;
                            if (len >= L) {
                                const uint32_t ci = ctx & ctx_mask;
                                const uint32_t gwd = __ldg(a.gate + (ci >> 5));
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
                    // a surviving beam whose parent labeling was just (re)created points at it
                    // again: the new beams publish hash + length, every orphan compares its parent
                    // hash with them
                    if (__any_sync(kFull, survive && plane < 0 && len > 0)) {
                        __syncwarp();  // this frame's readers of key / k32 / lanerank are done
                        if (take) {
                            sm.key[ford] = h;
                            sm.k32[ford] = (uint32_t)len | ((uint32_t)last << 30);  // len < 2^29 (arena limit)
                            sm.lanerank[ford] = (uint8_t)li;
                        }
                        int nmax = n_new;
#pragma unroll
                        for (int o = 16; o >= G; o >>= 1) {
                            const int x = __shfl_xor_sync(kFull, nmax, o);
                            nmax = x > nmax ? x : nmax;
                        }
                        __syncwarp();
                        for (int k = 0; k < nmax; ++k) {
                            const unsigned long long zh = sm.key[k];
                            const uint32_t zw = sm.k32[k];
                            const int zlen = (int)(zw & 0x3fffffffu);
                            if (k < n_new && survive && plane < 0 && len == zlen + 1 && hp == zh) {
                                plane = (int)sm.lanerank[k];
                                prep = ((int)(zw >> 30) == last) ? 1 : 0;
                            }
                        }
                    }
                    // the beam set changed: refresh which extensions are merged into a live child
                    sm.kill[li] = 0u;
                    __syncwarp();
                    if (run && alive && plane >= 0) reinterpret_cast<uint8_t *>(sm.kill)[plane * 4 + last] = 0x80;
                    __syncwarp();
                    if (run) km = alive ? (0x80808080u & ~sm.kill[li]) : 0u;
                    rmax = kNoRmax;  // the merge mask (or the beam) changed
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
                if (run && alive) sm.newlist[rank] = (uint8_t)li;
                __syncwarp();
                if (run) {
                    succ = (alive && rank + 1 < na) ? (int)sm.newlist[rank + 1] + gshift : lane;
                    first_lane = (int)sm.newlist[0] + gshift;
                    last_lane = (int)sm.newlist[na - 1] + gshift;
                }
                __syncwarp();
            }
        }
        if (live) t += nrun;

        // ------------------------------------------------------------ end of read
        const int succ_first = __shfl_sync(kFull, succ, first_lane);  // lane of the second best beam
        if (live && t >= T) {
            const long long seq_off = a.seq_offsets[read];
            const long long seq_cap = a.seq_offsets[read + 1] - seq_off;
            const double ln2 = 0.693147180559945309417;
            if (status == 0 && lane == first_lane) {
                const long long n = len;
                if (n > seq_cap) status = RADIAN_READ_SEQ_OVERFLOW;
                int c = node;
                for (long long i = n - 1; i >= 0; --i) {
                    const uint32_t w = arena[c] & 0x7fffffffu;
                    if (i < seq_cap) a.out_seq[seq_off + i] = (uint8_t)(w & 3u);
                    c = (int)(w >> 2);
                }
                a.out_len[read] = n;
                a.out_score[2 * read] = (ptot > 0.0) ? log(ptot) + (double)kacc * ln2 : -INFINITY;
                if (succ_first == first_lane) a.out_score[2 * read + 1] = NAN;
                a.out_status[read] = status;
                if (a.out_counters) {
                    a.out_counters[4 * read] = n_lookup;
                    a.out_counters[4 * read + 1] = n_combine;
                    a.out_counters[4 * read + 2] = n_tie;
                    a.out_counters[4 * read + 3] = 0;
                }
            }
            if (status == 0 && succ_first != first_lane && lane == succ_first)
                a.out_score[2 * read + 1] = (ptot > 0.0) ? log(ptot) + (double)kacc * ln2 : -INFINITY;
            if (status == RADIAN_READ_TRIE_OVERFLOW && li == 0) {
                a.out_len[read] = 0;
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
    // Small problems get the exact worst case; large ones assume >= 16 frames per base and report
    // RADIAN_READ_TRIE_OVERFLOW otherwise (the caller retries those reads with arena_nodes set).
    const int64_t G = group_size(beam_width);
    const int64_t exact = G * (max_frames + 1) + kNursery + 64;
    if (arena_nodes > 0) return arena_nodes < exact ? arena_nodes + kNursery : exact;
    if (exact <= (1 << 16)) return exact;
    int64_t cap = G * (max_frames / 16 + 64) + kNursery;
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

int decode_launch(const DecodeArgs &a, bool f64, int device, cudaStream_t stream)
{
    const bool lm = a.table != nullptr;
    DecodeLaunch dl;
    int rc = decode_pick(device, a.beam_width, lm, f64, a.out_counters != nullptr, a.ready != nullptr, &dl);
    if (rc) return rc;
    // no more groups than reads: extra CTAs would only touch the queue
    int64_t need = ((int64_t)a.n_reads + dl.groups_per_block - 1) / dl.groups_per_block;
    int grid = (int)(need < dl.grid ? need : dl.grid);
    if (grid < 1) grid = 1;
    DecodeArgs args = a;
    void *params[] = {(void *)&args};
    RADIAN_CUDA(cudaLaunchKernel(dl.kernel, dim3(grid), dim3(dl.block), params, dl.smem, stream));
    return 0;
}

}  // namespace radian
