// assemble.cuh -- K3: person assembly + final prune + COCO-ordered output.
//
// Replaces find_people (/root/reference/evaluate.py:279-498) and the tail of process() (:523-543).
//
// The reference is a state machine that consumes connections strictly in (limb, acceptance) order.  Two
// facts make it parallel inside a limb without changing any result:
//
//  (1) Ownership invariant.  At any time a peak id sits in at most one live row of `subset`: a new row is
//      only created when neither end point is in any row (:473), an assignment/replacement only happens
//      when exactly one row matched (:320), a merge moves ids between two disjoint rows (:403-424).  So the
//      scan "which rows hold idA in slot A or idB in slot B" (:304-318) returns at most one row per end
//      point (its third-match branch :314-316 is dead) and can be an O(1) lookup in an owner map
//      owner[part][peak] -> row.
//  (2) Independence.  A connection only reads and writes the rows it matched, plus a fresh row when it
//      matched none.  Connections of one limb have pairwise distinct end points (greedy matching, :265),
//      so if no row is matched by two connections of the limb they commute; the only order-dependent
//      quantity, the index of a newly created row, is a prefix count.
//
// One warp per image, one lane per connection of the current limb (chunks of 32).  Rounds: every pending
// connection looks up the rows it matches and stamps them with (round, lane); a connection is ELIGIBLE when it is
// the lowest-numbered pending connection on every row it matches -- then no earlier pending connection can touch
// (or re-home the end points of) those rows, so applying it now gives exactly the sequential result.  Eligible
// connections are pairwise row-disjoint and run their transition simultaneously; the rest wait for the next round
// (cross-person connection pairs (a,b'),(b,a') in crowds need 2-3 rounds, clean limbs need one).  All lanes run the
// same single-thread transition the sequential algorithm would.  Because rows are then created out of acceptance
// order, every row carries its birth stamp (limb, connection index): "row order" -- which decides j1 < j2 in a
// merge (:396) and the output order -- is birth order, and np.delete of a merged row (:424) is a tombstone.
// `subset` lives in shared memory as 32-byte row records and 16-byte (part, row) slot records.
//
// The loop is a chain of dependent shared-memory accesses on ONE warp: it runs at ~5 cycles per instruction whatever the
// rest of the SM does (clock trace: tools/trace_match_assemble.py), so its cost is its instruction count -- hence the
// record layout, the atomic-free eligibility test and the branch-free common transition below.  The prune + output phase
// runs on all warps of the CTA.
#pragma once

#include "common.cuh"

namespace spg {

struct AssembleArgs {
    int n_images, image_base, use_bulk;
    double len_rate, connection_tole, min_mean_score;
    int remove_recon, min_parts;
    int refresh_len_check;  // demo_image.py:414-415: the same-B refresh also checks the limb length
    // "records landed" signal folded into the kernel (spg_arm_wire_signal): the CTA that finishes last release-stores
    // wire_flag_value into *wire_flag (local or peer memory) -- no separate signalling kernel after the launch
    unsigned long long *wire_flag;
    unsigned long long wire_flag_value;
    unsigned int *done_counter;
    Workspace ws;
};

constexpr int kAssembleThreads = 32;

// `subset` in shared memory.  A row's scalar state is ONE 32-byte record and a (part, row) slot ONE 16-byte record, so that
// the per-connection transition -- a chain of dependent shared-memory accesses executed by a single warp, ~5 cycles per
// instruction -- costs two 16-byte loads per row and one per slot instead of a load plus address arithmetic per field.
struct alignas(16) RowRec {
    double total, maxlen;
    uint32_t mask;  // bit c set <=> slot c of the row holds a peak; id/sc of unset slots are the reference's -1 / -1.0
    int cnt, birth, alive;
};
struct alignas(16) SlotRec {
    double sc;
    int id, pad;
};

__host__ __device__ inline size_t align16(size_t b) { return (b + 15) & ~(size_t)15; }
__host__ __device__ inline size_t assemble_smem_bytes(int K, int capP, int capR) {
    return (size_t)capR * sizeof(RowRec) + (size_t)K * capR * sizeof(SlotRec) + align16((size_t)capR * sizeof(double))  // rows, slots, pscore
           + align16((size_t)K * capP * sizeof(float))                                                                 // peak scores
           + 2 * align16((size_t)capR * sizeof(int))                                                                   // postA, postB
           + align16((size_t)(K + 1) * sizeof(int))                                                                    // part offsets
           + align16((size_t)K * capP * sizeof(short));                                                                // owner
}

// connection tables of one image staged in shared memory (bulk copies): ij, score, norm [L][capP] + counts [L]
__host__ __device__ inline size_t assemble_conn_bytes(int L, int capP) {
    return align16((size_t)L * capP * (sizeof(uint32_t) + 2 * sizeof(double)) + (size_t)L * sizeof(int));
}

struct PersonTable {
    RowRec *row;      // [capR]
    SlotRec *slot;    // [K][capR]
    double *pscore;   // [capR] output phase
    float *ps;        // [K][capP] peak scores
    int *postA, *postB;  // [capR] round stamps of the limb loop (which lane reaches this row through its A / B end point); the
                         // output phase reuses postA as keep flag / output position
    int *off;         // [K + 1] part offsets
    short *owner;     // [K][capP] row that holds the peak, -1: none
    const double *px, *py;  // refined peak coordinates [K][capP]: global memory, or the fused kernel's shared-memory copy
    int K, capP, capR;
};

// carve the person table out of shared memory
__device__ __forceinline__ PersonTable make_person_table(unsigned char *base, int K, int capP, int capR) {
    PersonTable t;
    t.K = K; t.capP = capP; t.capR = capR;
    t.row = reinterpret_cast<RowRec *>(base);
    t.slot = reinterpret_cast<SlotRec *>(t.row + capR);
    t.pscore = reinterpret_cast<double *>(t.slot + (size_t)K * capR);
    unsigned char *p = reinterpret_cast<unsigned char *>(t.pscore) + align16((size_t)capR * sizeof(double));
    t.ps = reinterpret_cast<float *>(p);
    p += align16((size_t)K * capP * sizeof(float));
    t.postA = reinterpret_cast<int *>(p);
    p += align16((size_t)capR * sizeof(int));
    t.postB = reinterpret_cast<int *>(p);
    p += align16((size_t)capR * sizeof(int));
    t.off = reinterpret_cast<int *>(p);
    p += align16((size_t)(K + 1) * sizeof(int));
    t.owner = reinterpret_cast<short *>(p);
    t.px = t.py = nullptr;
    return t;
}

// Stamps, owner map and part offsets of image n, by one warp (shared-memory stores + K counters from global memory).
__device__ __forceinline__ void init_person_rows(const PersonTable &t, const Workspace &ws, int n, int lane) {
    const int K = t.K, capP = t.capP, capR = t.capR;
    // part offsets: exclusive prefix sum of the (capped) peak counts, K <= 32
    const int cnt = lane < K ? min(ws.peak_count[(size_t)n * K + lane], capP) : 0;
    int inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += v;
    }
    if (lane < K) t.off[lane] = inc - cnt;
    if (lane == K - 1) t.off[K] = inc;
    for (int i = lane; i < K * capP; i += 32) t.owner[i] = -1;
    for (int i = lane; i < capR; i += 32) {
        t.postA[i] = 0;
        t.postB[i] = 0;
    }
}

// Peak scores (and, s_xy != nullptr, the refined coordinates) of image n into shared memory, by `nthreads` cooperating
// threads.  The limb loop reads the scores; the output phase gathers 17 coordinates per person (from L2 that was 20 % of
// the first stand-alone kernel's time) and the fused kernel's matchers take the limb lengths from them.
__device__ __forceinline__ void stage_peaks(const PersonTable &t, const Workspace &ws, int n, int tid, int nthreads, double *s_xy) {
    const int KP = t.K * t.capP;
    const size_t g = (size_t)n * KP;
    for (int i = tid; i < KP; i += nthreads) {
        t.ps[i] = ws.peak_score[g + i];
        if (s_xy) {
            s_xy[i] = ws.peak_x[g + i];
            s_xy[KP + i] = ws.peak_y[g + i];
        }
    }
}

__device__ __forceinline__ RowRec load_row(const RowRec *r) {
    RowRec v;
    const double2 d = *reinterpret_cast<const double2 *>(r);
    const int4 q = reinterpret_cast<const int4 *>(r)[1];
    v.total = d.x; v.maxlen = d.y; v.mask = (uint32_t)q.x; v.cnt = q.y; v.birth = q.z; v.alive = q.w;
    return v;
}
__device__ __forceinline__ void store_row_head(RowRec *r, double total, double maxlen) {
    *reinterpret_cast<double2 *>(r) = make_double2(total, maxlen);
}
__device__ __forceinline__ void store_row_tail(RowRec *r, uint32_t mask, int cnt, int birth, int alive) {
    reinterpret_cast<int4 *>(r)[1] = make_int4((int)mask, cnt, birth, alive);
}
__device__ __forceinline__ SlotRec load_slot(const SlotRec *p) {
    const int4 v = *reinterpret_cast<const int4 *>(p);
    SlotRec s;
    s.sc = __hiloint2double(v.y, v.x);
    s.id = v.z; s.pad = 0;
    return s;
}
__device__ __forceinline__ void store_slot(SlotRec *p, double sc, int id) {
    *reinterpret_cast<int4 *>(p) = make_int4(__double2loint(sc), __double2hiint(sc), id, 0);
}

// The reference's transition for a connection that matched TWO different rows (:385-460), executed by one thread.
__device__ __forceinline__ void apply_two_rows(const PersonTable &t, const AssembleArgs &a, int A, int B, int ra, int rb, double s, double len) {
    const int capP = t.capP, capR = t.capR;
    RowRec r_a = load_row(t.row + ra), r_b = load_row(t.row + rb);
    // j1 before j2 in row (= birth) order
    const bool a_first = r_a.birth < r_b.birth;
    const int j1 = a_first ? ra : rb, j2 = a_first ? rb : ra;
    const RowRec r1 = a_first ? r_a : r_b, r2 = a_first ? r_b : r_a;
    const uint32_t m1 = r1.mask, m2 = r2.mask;
    if ((m1 & m2) == 0u) {  // disjoint -> merge j2 into j1 (:403-424)
        double m = INFINITY;  // min over the connection scores present in either row (:405-407)
        for (uint32_t b = m1; b; b &= b - 1) m = fmin(m, t.slot[(__ffs(b) - 1) * capR + j1].sc);
        for (uint32_t b = m2; b; b &= b - 1) m = fmin(m, t.slot[(__ffs(b) - 1) * capR + j2].sc);
        const double ml1 = r1.maxlen;
        if (s < __dmul_rn(a.connection_tole, m) || __dmul_rn(a.len_rate, ml1) <= len) return;
        // the "+1" trick (:415) on both columns.  Slots absent from j2 add (-1 + 1) = 0 to j1: unchanged.
        // Slots present in j2 are absent from j1 (disjoint): id -1 + (id2 + 1), score -1.0 + (sc2 + 1.0).
        for (uint32_t b = m2; b; b &= b - 1) {
            const int c = __ffs(b) - 1;
            const SlotRec s2 = load_slot(t.slot + c * capR + j2);
            store_slot(t.slot + c * capR + j1, __dadd_rn(-1.0, __dadd_rn(s2.sc, 1.0)), s2.id);
            t.owner[c * capP + (s2.id - t.off[c])] = (short)j1;
        }
        store_row_head(t.row + j1, __dadd_rn(__dadd_rn(r1.total, r2.total), s), len > ml1 ? len : ml1);  // :419-422: keeps j1's own longest limb
        store_row_tail(t.row + j1, m1 | m2, r1.cnt + r2.cnt, r1.birth, 1);
        t.row[j2].alive = 0;  // np.delete(subset, j2) (:424)
        return;
    }
    // overlapping rows (:426-460): only remove_recon > 0 has side effects
    if (a.remove_recon <= 0) return;    // (the lookups below cannot fail: a peak id sits in exactly one slot of one row)
    const bool a_in_1 = (ra == j1);     // idA is in j1 iff j1 is the row that owns it
    const int c1 = a_in_1 ? A : B, c2 = a_in_1 ? B : A;
    const double e1 = t.slot[c1 * capR + j1].sc, e2 = t.slot[c2 * capR + j2].sc;
    if (s < e1 && s < e2) return;
    int small_j = j1, rc = c1;
    if (e1 > e2) { small_j = j2; rc = c2; }
    const SlotRec sr = load_slot(t.slot + rc * capR + small_j);
    const int ridx = sr.id - t.off[rc];
    RowRec *rs = t.row + small_j;
    rs->total = __dsub_rn(rs->total, __dadd_rn((double)t.ps[rc * capP + ridx], sr.sc));
    rs->mask &= ~(1u << rc);
    rs->cnt -= 1;
    t.owner[rc * capP + ridx] = -1;
}

// The transition for a connection that matched NO row (new person, :473-488) or exactly ONE (:320-383: assign, replace or
// refresh slot B of the matched row).  Written without data-dependent branches -- every lane loads its row and slot
// records, decides with predicates and stores under them -- because these two cases are what a round of the limb loop
// almost always consists of.
struct ConnCtx {
    int A, B, ia, jb, idA, idB, offB;
    double s, len, psA, psB;
    SlotRec *slotA, *slotB;   // column bases: slot + part * capR
    short *ownA, *ownB;       // this connection's two owner entries
    short *ownB_base;         // owner + B * capP
    const float *psB_base;    // ps + B * capP
};
__device__ __forceinline__ void apply_zero_or_one(const PersonTable &t, const AssembleArgs &a, const ConnCtx &c, int ra, int rb, int new_row,
                                                  int birth) {
    const bool isnew = ra < 0 && rb < 0;
    const int j = isnew ? new_row : (ra >= 0 ? ra : rb);
    RowRec *rp = t.row + j;
    RowRec r;
    r.total = 0.0; r.maxlen = 0.0; r.mask = 0u; r.cnt = 0; r.birth = birth; r.alive = 1;
    if (!isnew) r = load_row(rp);
    const bool hasB = (r.mask >> c.B) & 1u;
    SlotRec *sbp = c.slotB + j;
    SlotRec sb;
    sb.sc = -1.0; sb.id = -1; sb.pad = 0;
    if (hasB) sb = load_slot(sbp);
    const int oldIdx = hasB ? sb.id - c.offB : c.jb;
    const double psOld = (double)c.psB_base[oldIdx];
    const double reach = __dmul_rn(a.len_rate, r.maxlen);
    const double add = __dadd_rn(c.psB, c.s);
    const bool assign = !isnew && !hasB && reach > c.len;                                                        // :323-342
    const bool replace = !isnew && !assign && sb.id != c.idB && hasB && !(sb.sc >= c.s) && !(reach <= c.len);     // :346-363
    const bool refresh = !isnew && !assign && sb.id == c.idB && sb.sc <= c.s && !(a.refresh_len_check && reach <= c.len);  // :368-380
    // builtin sum() of the two end-point scores, then + s (:484) | total + add | (total - (old end point + old score)) + add
    const double tot_new = __dadd_rn(__dadd_rn(__dadd_rn(0.0, c.psA), c.psB), c.s);
    const double tot_swap = __dadd_rn(__dsub_rn(r.total, __dadd_rn(psOld, sb.sc)), add);
    const double tot = isnew ? tot_new : (assign ? __dadd_rn(r.total, add) : tot_swap);
    if (isnew | assign | replace | refresh) {
        store_row_head(rp, tot, (isnew || c.len > r.maxlen) ? c.len : r.maxlen);
        store_slot(sbp, c.s, c.idB);  // (refresh: the id is idB already)
    }
    if (isnew | assign | replace) {
        if (replace) c.ownB_base[oldIdx] = -1;
        *c.ownB = (short)j;
    }
    if (isnew | assign) store_row_tail(rp, r.mask | (1u << c.B) | (isnew ? (1u << c.A) : 0u), r.cnt + (isnew ? 2 : 1), r.birth, r.alive);
    if (isnew) {
        store_slot(c.slotA + j, c.s, c.idA);
        *c.ownA = (short)j;
    }
}

struct AsmResult {  // what the limb loop hands to the output phase
    int nrows;
    uint32_t flags;
};

// The limb loop of find_people, by ONE warp.  `t` must be initialised (stamps, owner map, offsets, peak scores).  FUSED:
// matcher warps of the same CTA write each limb's rows into the shared-memory tables and raise s_ready[k]; limb k's flag is
// acquired right before the limb is consumed, so matching limbs k+1.. overlaps assembling limb k.
//
// Which connections of a round may run together: a connection reaches a row through its A end point (the row that owns the
// A peak) and/or through its B end point.  The A peaks of a limb's connections are pairwise different, and a row holds one
// peak per part, so no two connections reach the same row through A -- likewise through B.  A row is therefore reached by at
// most two pending connections, one through A and one through B, and "lowest pending connection on every row it reaches"
// needs no atomics: every pending lane stamps postA[ra] and postB[rb] with (round, lane) and looks up postB[ra] and postA[rb].
template <bool FUSED>
__device__ __forceinline__ AsmResult assemble_limbs(const AssembleArgs &a, const PersonTable &t, const double *s_cs, const double *s_cn,
                                                    const uint32_t *s_cij, const int *s_cc, int lane, const int *s_ready) {
    const Workspace &ws = a.ws;
    const int L = ws.L, capP = ws.capP, capR = ws.capR;
    int nrows = 0, stamp = 0;
    uint32_t flags = 0;
    bool overflow = false;
    for (int k = 0; k < L && !overflow; k++) {
        SPG_TR(160 + 4 * k, 0);
        int trace_rounds = 0;
        (void)trace_rounds;
        if (FUSED) {  // acquire: limb k's rows and counter are in shared memory
            int r;
            do {
                asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(r) : "r"(smem_u32(s_ready + k)) : "memory");
                if (!r) __nanosleep(20);
            } while (!r);
        }
        const int cc = s_cc[k];
        SPG_TR(160 + 4 * k + 1, cc);
        if (cc < 0) continue;  // special_k (:290)
        ConnCtx c;
        c.A = ws.limbs[2 * k]; c.B = ws.limbs[2 * k + 1];
        const int offA = t.off[c.A];
        c.offB = t.off[c.B];
        c.slotA = t.slot + c.A * capR;
        c.slotB = t.slot + c.B * capR;
        c.ownB_base = t.owner + c.B * capP;
        c.psB_base = t.ps + c.B * capP;
        for (int chunk = 0; chunk < cc && !overflow; chunk += 32) {
            const int mine_c = min(chunk + lane, cc - 1);
            const uint32_t my_ij = s_cij[k * capP + mine_c];
            c.s = s_cs[k * capP + mine_c];
            c.len = s_cn[k * capP + mine_c];
            const int in_chunk = min(32, cc - chunk);
            c.ia = (int)(my_ij >> 16); c.jb = (int)(my_ij & 0xffff);
            c.idA = offA + c.ia; c.idB = c.offB + c.jb;
            c.psA = (double)t.ps[c.A * capP + c.ia];
            c.psB = (double)c.psB_base[c.jb];
            c.ownA = t.owner + c.A * capP + c.ia;
            c.ownB = c.ownB_base + c.jb;
            const int birth = (k << 8) | (chunk + lane);
            uint32_t pending = in_chunk == 32 ? 0xffffffffu : ((1u << in_chunk) - 1u);
            while (pending) {
                stamp++;
                const bool mine = (pending >> lane) & 1u;
                int ra = -1, rb = -1;
                const int tag = (stamp << 5) | lane;
                if (mine) {
                    ra = *c.ownA;
                    rb = *c.ownB;
                    if (ra >= 0) t.postA[ra] = tag;
                    if (rb >= 0) t.postB[rb] = tag;
                }
                __syncwarp();
                bool eligible = mine;
                if (ra >= 0) {  // a lower pending lane reaches my A row through its B end point
                    const int y = t.postB[ra];
                    if ((y >> 5) == stamp && (y & 31) < lane) eligible = false;
                }
                if (rb >= 0) {
                    const int z = t.postA[rb];
                    if ((z >> 5) == stamp && (z & 31) < lane) eligible = false;
                }
                const bool two = ra >= 0 && rb >= 0 && ra != rb;
                const uint32_t emask = __ballot_sync(0xffffffffu, eligible);
                const uint32_t creates = __ballot_sync(0xffffffffu, eligible && ra < 0 && rb < 0);
                const uint32_t twos = __ballot_sync(0xffffffffu, eligible && two);
                if (nrows + __popc(creates) > capR) {
                    flags |= kStRowOverflow;
                    overflow = true;
                    break;
                }
                // eligible connections are pairwise row-disjoint: the two forms may run in either order
                if (eligible && !two) apply_zero_or_one(t, a, c, ra, rb, nrows + __popc(creates & ((1u << lane) - 1u)), birth);
                if (twos) {  // warp-uniform: most rounds have none
                    if (eligible && two) apply_two_rows(t, a, c.A, c.B, ra, rb, c.s, c.len);
                }
                nrows += __popc(creates);
                pending &= ~emask;
                __syncwarp();
                trace_rounds++;
            }
        }
        SPG_TR(160 + 4 * k + 2, nrows);
        SPG_TRV(160 + 4 * k + 3, trace_rounds * 256 + cc);
    }
    SPG_TR(300, nrows);
    AsmResult r;
    r.nrows = nrows;
    r.flags = __reduce_or_sync(0xffffffffu, flags);
    return r;
}

// Prune (:491-496) + outputs of one image by `nthreads` = 32 * nwarps cooperating threads (all warps of the CTA).  Kept rows
// keep their relative (birth) order.  `s_stage`: shared memory that is dead by now (the connection tables) -- the wire
// record's rows are staged there and leave in one coalesced copy, so a record in a peer GPU's memory costs a few 128-byte
// NVLink writes per image.
__device__ __forceinline__ void emit_people(const AssembleArgs &a, const PersonTable &t, double *s_stage, size_t stage_bytes, int *s_out, int n,
                                            int img_in_call, AsmResult res, int tid, int nthreads) {
    const Workspace &ws = a.ws;
    const int K = ws.K, capP = ws.capP, capR = ws.capR, J = ws.J;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
    const int nrows = res.nrows;
    uint32_t flags = res.flags;
    const int RS = K + 2;
    double *g_subset = ws.subset + (size_t)n * capR * RS * 2;
    double *g_xy = ws.people_xy + (size_t)n * capR * J * 2;
    double *g_score = ws.people_score + (size_t)n * capR;
    const int WR = 2 * J + 2;  // x,y per joint, person score, presence mask
    const bool wire_on = ws.wire != nullptr && (size_t)min(ws.wire_rows, capR) * WR * sizeof(double) <= stage_bytes;
    double *s_wire = s_stage;
    int *keep = t.postA;      // 1: the row survives the prune
    int *keep_pos = t.postB;  // 0: dropped, p + 2: kept at output position p
    // keep flags and person scores: one row per thread (the two float64 divisions are ~100 instructions each)
    for (int j = tid; j < nrows; j += nthreads) {
        const RowRec r = load_row(t.row + j);
        bool kp = false;
        if (r.alive) kp = !(r.cnt < a.min_parts || __ddiv_rn(r.total, (double)r.cnt) < a.min_mean_score);
        keep[j] = kp ? 1 : 0;
        if (kp) t.pscore[j] = __dsub_rn(1.0, __ddiv_rn(1.0, r.total));  // :541
    }
    if (tid == 0) *s_out = 0;
    __syncthreads();
    // A kept row's output position = number of kept rows born earlier: four threads per row, a quarter of the rows each.
    const int rows_per_pass = nthreads >> 2;
    for (int base = 0; base < nrows; base += rows_per_pass) {
        const int j = base + (tid >> 2), q = tid & 3;
        const bool act = j < nrows;
        const int mine = act ? t.row[j].birth : 0;
        int o = 0, kept = 0;
        if (act)
            for (int u = q; u < nrows; u += 4) {
                const bool ku = keep[u] != 0;
                kept += ku ? 1 : 0;
                o += (ku && t.row[u].birth < mine) ? 1 : 0;
            }
        o += __shfl_xor_sync(0xffffffffu, o, 1);
        o += __shfl_xor_sync(0xffffffffu, o, 2);
        kept += __shfl_xor_sync(0xffffffffu, kept, 1);
        kept += __shfl_xor_sync(0xffffffffu, kept, 2);
        if (act && q == 0) {
            keep_pos[j] = keep[j] != 0 ? o + 2 : 0;
            if (j == 0) *s_out = kept;
        }
    }
    __syncthreads();
    const int out = *s_out;
    // one thread per (row, column): `subset` rows (K + 2 columns of two doubles), then the joints (J columns), a group of
    // nthreads / columns rows per pass -- the thread's column and row-in-group are fixed, so a pass is a flag lookup,
    // two 16-byte loads and one 16-byte store
    {
        // (K + 2 > nthreads -- 32 parts in the one-warp stand-alone kernel -- : two column passes per row)
        const int cols = RS, rows_pp = max(nthreads / cols, 1);
        const int jr = tid / cols;
        for (int col = tid - jr * cols; col < cols; col += nthreads)
            for (int j = jr; j < nrows && jr < rows_pp; j += rows_pp) {
                const int tj = keep_pos[j];
                if (tj == 0) continue;
                const RowRec r = load_row(t.row + j);
                double2 v;
                if (col < K) {
                    SlotRec sl;
                    sl.sc = -1.0; sl.id = -1; sl.pad = 0;
                    if ((r.mask >> col) & 1u) sl = load_slot(t.slot + col * capR + j);
                    v = make_double2((double)sl.id, sl.sc);
                } else if (col == K) {
                    v = make_double2(r.total, -1.0);
                } else {
                    v = make_double2((double)r.cnt, r.maxlen);
                }
                reinterpret_cast<double2 *>(g_subset)[(size_t)(tj - 2) * RS + col] = v;
            }
    }
    if (J > 0) {
        const int rows_pp = nthreads / J;
        const int jr = tid / J, g = tid - jr * J;
        if (jr < rows_pp) {
            const int part = ws.out_from_part[g];
            const int offp = t.off[part];
            for (int j = jr; j < nrows; j += rows_pp) {
                const int tj = keep_pos[j];
                if (tj == 0) continue;
                const int o = tj - 2;
                double x = 0.0, y = 0.0;  // :523-539
                if ((t.row[j].mask >> part) & 1u) {
                    const int idx = t.slot[part * capR + j].id - offp;
                    x = t.px[part * capP + idx];
                    y = t.py[part * capP + idx];
                }
                reinterpret_cast<double2 *>(g_xy)[(size_t)o * J + g] = make_double2(x, y);
                if (wire_on && o < ws.wire_rows) {
                    s_wire[(size_t)o * WR + 2 * g + 0] = x;
                    s_wire[(size_t)o * WR + 2 * g + 1] = y;
                }
            }
        }
    }
    for (int j = tid; j < nrows; j += nthreads) {  // person score and presence mask
        const int tj = keep_pos[j];
        if (tj == 0) continue;
        const int o = tj - 2;
        const double pscore = t.pscore[j];
        g_score[o] = pscore;
        if (wire_on && o < ws.wire_rows) {
            const uint32_t mj = t.row[j].mask;
            unsigned long long pm = 0ull;
            for (int g = 0; g < J; g++) pm |= (unsigned long long)((mj >> ws.out_from_part[g]) & 1u) << g;
            s_wire[(size_t)o * WR + 2 * J] = pscore;
            reinterpret_cast<unsigned long long *>(s_wire)[(size_t)o * WR + 2 * J + 1] = pm;
        }
    }
    if (ws.wire != nullptr && (!wire_on || out > ws.wire_rows)) flags |= kStWireOverflow;
    __shared__ uint32_t s_status;
    if (tid == 0) {
        ws.n_persons[n] = out;
        s_status = flags ? (atomicOr(&ws.status[n], flags) | flags) : ws.status[n];
    }
    __syncthreads();
    if (ws.wire != nullptr) {
        const size_t rec_bytes = 8 + (size_t)ws.wire_rows * WR * sizeof(double);
        unsigned char *rec = ws.wire + (size_t)(ws.wire_first + (long long)img_in_call) * rec_bytes;
        const int wn = wire_on ? min(out, ws.wire_rows) : 0;
        if (tid == 0) {
            reinterpret_cast<int *>(rec)[0] = wn;
            reinterpret_cast<uint32_t *>(rec)[1] = s_status;
        }
        double *rows = reinterpret_cast<double *>(rec + 8);
        for (int i = tid; i < wn * WR; i += nthreads) rows[i] = s_wire[i];
    }
    if (a.wire_flag != nullptr) {  // last CTA done publishes the step (the threadFenceReduction pattern, system scope)
        __syncthreads();           // every thread's record stores are ordered before thread 0's fence
        if (tid == 0) {
            __threadfence_system();
            const unsigned int prev = atomicAdd(a.done_counter, 1u);
            if (prev == (unsigned int)a.n_images - 1u) {
                *a.done_counter = 0u;  // re-armed for the next launch (stream order separates the launches)
                __threadfence_system();
                asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.wire_flag), "l"(a.wire_flag_value) : "memory");
            }
        }
    }
    SPG_TR(301, 0);
}

// The stand-alone kernel: one warp per image; the image's connection tables are fetched from global memory up front (by
// the bulk-copy engine) so that the serial limb loop never waits on L2.
__device__ __forceinline__ void assemble_image(const AssembleArgs &a, unsigned char *smem_raw, uint64_t &bar, int *s_out, int n, int img_in_call,
                                               int lane) {
    const Workspace &ws = a.ws;
    const int K = ws.K, L = ws.L, capP = ws.capP, capR = ws.capR;
    // shared memory: [conn_score | conn_norm | conn_ij | conn_count] of this image, then the person table
    const size_t LC = (size_t)L * capP;
    double *s_cs = reinterpret_cast<double *>(smem_raw);
    double *s_cn = s_cs + LC;
    uint32_t *s_cij = reinterpret_cast<uint32_t *>(s_cn + LC);
    int *s_cc = reinterpret_cast<int *>(s_cij + LC);
    PersonTable t = make_person_table(smem_raw + assemble_conn_bytes(L, capP), K, capP, capR);
    t.px = ws.peak_x + (size_t)n * K * capP;
    t.py = ws.peak_y + (size_t)n * K * capP;
    const size_t img_conn = (size_t)n * LC;
    if (a.use_bulk) {
        if (lane == 0) {
            mbar_init(&bar, 1);
            fence_mbar_init();
            mbar_expect_tx(&bar, (uint32_t)(LC * (sizeof(uint32_t) + 2 * sizeof(double))));
            bulk_g2s(s_cs, ws.conn_score + img_conn, (uint32_t)(LC * sizeof(double)), &bar);
            bulk_g2s(s_cn, ws.conn_norm + img_conn, (uint32_t)(LC * sizeof(double)), &bar);
            bulk_g2s(s_cij, ws.conn_ij + img_conn, (uint32_t)(LC * sizeof(uint32_t)), &bar);
        }
    } else {
        for (size_t i = lane; i < LC; i += 32) {
            s_cs[i] = ws.conn_score[img_conn + i];
            s_cn[i] = ws.conn_norm[img_conn + i];
            s_cij[i] = ws.conn_ij[img_conn + i];
        }
    }
    for (int k = lane; k < L; k += 32) s_cc[k] = ws.conn_count[(size_t)n * L + k];
    init_person_rows(t, ws, n, lane);
    stage_peaks(t, ws, n, lane, 32, nullptr);
    __syncwarp();
    if (a.use_bulk) mbar_wait(&bar, 0);
    const AsmResult res = assemble_limbs<false>(a, t, s_cs, s_cn, s_cij, s_cc, lane, nullptr);
    __syncthreads();
    emit_people(a, t, reinterpret_cast<double *>(smem_raw), assemble_conn_bytes(L, capP), s_out, n, img_in_call, res, lane, 32);
}

__global__ void __launch_bounds__(kAssembleThreads) assemble_kernel(AssembleArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ int s_out;
    if ((int)blockIdx.x >= a.n_images) return;
    assemble_image(a, smem_raw, bar, &s_out, a.image_base + blockIdx.x, blockIdx.x, threadIdx.x);
}

}  // namespace spg
