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
// connection looks up the rows it matches and posts its lane index on them with a shared-memory atomicMin;
// a connection is ELIGIBLE when it is the lowest-numbered pending connection on every row it matches --
// then no earlier pending connection can touch (or re-home the end points of) those rows, so applying it now
// gives exactly the sequential result.  Eligible connections are pairwise row-disjoint and run their
// transition simultaneously; the rest wait for the next round (cross-person connection pairs (a,b'),(b,a')
// in crowds need 2-3 rounds, clean limbs need one).  All lanes run the same single-thread transition
// function the sequential algorithm would.  Because rows are then created out of acceptance order, every
// row carries its birth stamp (limb, connection index): "row order" -- which decides j1 < j2 in a merge
// (:396) and the output order -- is birth order, and np.delete of a merged row (:424) is a tombstone.
// `subset` lives in shared memory as structure-of-arrays.
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

__host__ __device__ inline size_t assemble_smem_bytes(int K, int capP, int capR) {
    size_t b = (size_t)K * capR * sizeof(double)      // sc
               + 2 * (size_t)capR * sizeof(double)    // total, maxlen
               + (size_t)K * capR * sizeof(int)       // id
               + 4 * (size_t)capR * sizeof(int)       // cnt, touch, birth, mask
               + (size_t)K * capP * sizeof(float)     // peak scores
               + (size_t)(K + 1) * sizeof(int)        // part offsets
               + (size_t)K * capP * sizeof(short)     // owner
               + (size_t)capR;                        // alive
    return (b + 15) & ~(size_t)15;
}

// connection tables of one image staged in shared memory (bulk copies): ij, score, norm [L][capP] + counts [L]
__host__ __device__ inline size_t assemble_conn_bytes(int L, int capP) {
    return (size_t)L * capP * (sizeof(uint32_t) + 2 * sizeof(double)) + (((size_t)L * sizeof(int) + 15) & ~(size_t)15);
}

struct PersonTable {
    double *sc, *total, *maxlen;
    int *id, *cnt, *touch, *birth, *off;
    uint32_t *mask;  // bit c set <=> slot c of the row holds a peak; id/sc of unset slots are the reference's -1 / -1.0
    float *ps;
    short *owner;
    unsigned char *alive;
    const double *px, *py;  // refined peak coordinates [K][capP]: global memory, or the fused kernel's shared-memory copy
    int K, capP, capR;
};

// carve the person table out of shared memory
__device__ __forceinline__ PersonTable make_person_table(unsigned char *table_base, int K, int capP, int capR) {
    PersonTable t;
    t.K = K; t.capP = capP; t.capR = capR;
    t.sc = reinterpret_cast<double *>(table_base);
    t.total = t.sc + (size_t)K * capR;
    t.maxlen = t.total + capR;
    t.id = reinterpret_cast<int *>(t.maxlen + capR);
    t.cnt = t.id + (size_t)K * capR;
    t.touch = t.cnt + capR;
    t.birth = t.touch + capR;
    t.mask = reinterpret_cast<uint32_t *>(t.birth + capR);
    t.ps = reinterpret_cast<float *>(t.mask + capR);
    t.off = reinterpret_cast<int *>(t.ps + (size_t)K * capP);
    t.owner = reinterpret_cast<short *>(t.off + (K + 1));
    t.alive = reinterpret_cast<unsigned char *>(t.owner + (size_t)K * capP);
    t.px = t.py = nullptr;
    return t;
}

// peak scores, owner map, row flags and part offsets of image n, by `nthreads` cooperating threads (one warp in the
// stand-alone kernel, the whole CTA in the fused one); s_xy != nullptr: also stage the refined coordinates (the output
// phase gathers 17 of them per person -- from L2 that was 20 % of the stand-alone kernel's time)
__device__ __forceinline__ void init_person_table(const PersonTable &t, const Workspace &ws, int n, int tid, int nthreads, double *s_xy) {
    const int K = t.K, capP = t.capP, capR = t.capR;
    if (tid == 0) {
        int acc = 0;
        for (int c = 0; c < K; c++) {
            t.off[c] = acc;
            acc += min(ws.peak_count[(size_t)n * K + c], capP);
        }
        t.off[K] = acc;
    }
    for (int i = tid; i < K * capP; i += nthreads) {
        t.ps[i] = ws.peak_score[(size_t)n * K * capP + i];
        t.owner[i] = -1;
        if (s_xy) {
            s_xy[i] = ws.peak_x[(size_t)n * K * capP + i];
            s_xy[(size_t)K * capP + i] = ws.peak_y[(size_t)n * K * capP + i];
        }
    }
    for (int i = tid; i < capR; i += nthreads) {
        t.touch[i] = 0x7fffffff;
        t.alive[i] = 0;
    }
}

// The reference's per-connection transition (:320-488), executed by ONE thread.  `new_row` is the row index to
// use if the connection matches nothing, `birth` its stamp.  Returns status flags.
__device__ __forceinline__ uint32_t apply_connection(const PersonTable &t, const AssembleArgs &a, int A, int B, int ia,
                                                     int jb, double s, double len, int new_row, int birth) {
    const int capP = t.capP, capR = t.capR;
    const int idA = t.off[A] + ia, idB = t.off[B] + jb;
    const int ra = t.owner[A * capP + ia], rb = t.owner[B * capP + jb];
    if (ra < 0 && rb < 0) {  // new person (:473-488)
        const int j = new_row;
        t.mask[j] = (1u << A) | (1u << B);
        t.id[A * capR + j] = idA;
        t.sc[A * capR + j] = s;
        t.id[B * capR + j] = idB;
        t.sc[B * capR + j] = s;
        t.cnt[j] = 2;
        t.maxlen[j] = len;
        // builtin sum() of the two end-point scores, then + s (:484)
        t.total[j] = __dadd_rn(__dadd_rn(__dadd_rn(0.0, (double)t.ps[A * capP + ia]), (double)t.ps[B * capP + jb]), s);
        t.alive[j] = 1;
        t.birth[j] = birth;
        t.owner[A * capP + ia] = (short)j;
        t.owner[B * capP + jb] = (short)j;
        return 0;
    }
    if (ra >= 0 && rb >= 0 && ra != rb) {  // two rows (:385-460), j1 before j2 in row (= birth) order
        const bool a_first = t.birth[ra] < t.birth[rb];
        const int j1 = a_first ? ra : rb, j2 = a_first ? rb : ra;
        const uint32_t m1 = t.mask[j1], m2 = t.mask[j2];
        if ((m1 & m2) == 0u) {  // disjoint -> merge j2 into j1 (:403-424)
            double m = INFINITY;  // min over the connection scores present in either row (:405-407)
            for (uint32_t b = m1; b; b &= b - 1) m = fmin(m, t.sc[(__ffs(b) - 1) * capR + j1]);
            for (uint32_t b = m2; b; b &= b - 1) m = fmin(m, t.sc[(__ffs(b) - 1) * capR + j2]);
            const double ml1 = t.maxlen[j1];
            if (s < __dmul_rn(a.connection_tole, m) || __dmul_rn(a.len_rate, ml1) <= len) return 0;
            // the "+1" trick (:415) on both columns.  Slots absent from j2 add (-1 + 1) = 0 to j1: unchanged.
            // Slots present in j2 are absent from j1 (disjoint): id -1 + (id2 + 1), score -1.0 + (sc2 + 1.0).
            for (uint32_t b = m2; b; b &= b - 1) {
                const int c = __ffs(b) - 1;
                const int i2 = t.id[c * capR + j2];
                t.id[c * capR + j1] = i2;
                t.sc[c * capR + j1] = __dadd_rn(-1.0, __dadd_rn(t.sc[c * capR + j2], 1.0));
                t.owner[c * capP + (i2 - t.off[c])] = (short)j1;
            }
            t.mask[j1] = m1 | m2;
            t.total[j1] = __dadd_rn(__dadd_rn(t.total[j1], t.total[j2]), s);  // :419, :421
            t.cnt[j1] += t.cnt[j2];
            t.maxlen[j1] = len > ml1 ? len : ml1;  // keeps j1's own longest limb (:422)
            t.alive[j2] = 0;                        // np.delete(subset, j2) (:424)
            return 0;
        }
        // overlapping rows (:426-460): only remove_recon > 0 has side effects
        if (a.remove_recon <= 0) return 0;  // (the lookups below cannot fail: a peak id sits in exactly one slot of one row)
        const bool a_in_1 = (ra == j1);     // idA is in j1 iff j1 is the row that owns it
        const int c1 = a_in_1 ? A : B, c2 = a_in_1 ? B : A;
        const double e1 = t.sc[c1 * capR + j1], e2 = t.sc[c2 * capR + j2];
        if (s < e1 && s < e2) return 0;
        int small_j = j1, rc = c1;
        if (e1 > e2) { small_j = j2; rc = c2; }
        const int rid = t.id[rc * capR + small_j];
        const int ridx = rid - t.off[rc];
        t.total[small_j] = __dsub_rn(t.total[small_j], __dadd_rn((double)t.ps[rc * capP + ridx], t.sc[rc * capR + small_j]));
        t.mask[small_j] &= ~(1u << rc);
        t.cnt[small_j] -= 1;
        t.owner[rc * capP + ridx] = -1;
        return 0;
    }
    // exactly one row (:320-383) -- always slot B of the matched row
    const int j = ra >= 0 ? ra : rb;
    const uint32_t mj = t.mask[j];
    const bool hasB = (mj >> B) & 1u;
    const int oldB = hasB ? t.id[B * capR + j] : -1;
    const double scB = hasB ? t.sc[B * capR + j] : -1.0;
    const double ml = t.maxlen[j];
    const double reach = __dmul_rn(a.len_rate, ml);
    const double add = __dadd_rn((double)t.ps[B * capP + jb], s);
    if (!hasB && reach > len) {  // assign (:323-342)
        t.mask[j] = mj | (1u << B);
        t.id[B * capR + j] = idB;
        t.sc[B * capR + j] = s;
        t.cnt[j] += 1;
        t.total[j] = __dadd_rn(t.total[j], add);
        t.maxlen[j] = len > ml ? len : ml;
        t.owner[B * capP + jb] = (short)j;
    } else if (oldB != idB) {
        if (hasB && !(scB >= s) && !(reach <= len)) {  // replace (:346-363); an empty slot only gets here when too long
            const int oldIdx = oldB - t.off[B];
            const double sub = __dadd_rn((double)t.ps[B * capP + oldIdx], scB);
            t.total[j] = __dadd_rn(__dsub_rn(t.total[j], sub), add);
            t.id[B * capR + j] = idB;
            t.sc[B * capR + j] = s;
            t.maxlen[j] = len > ml ? len : ml;
            t.owner[B * capP + oldIdx] = -1;
            t.owner[B * capP + jb] = (short)j;
        }
    } else if (scB <= s) {  // same B, refresh its score (:368-380)
        if (a.refresh_len_check && reach <= len) return 0;  // demo_image.py:414-415 only
        const double sub = __dadd_rn((double)t.ps[B * capP + jb], scB);
        t.total[j] = __dadd_rn(__dsub_rn(t.total[j], sub), add);
        t.sc[B * capR + j] = s;
        t.maxlen[j] = len > ml ? len : ml;
    }
    return 0;
}

__device__ __forceinline__ double shfl_f64(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// One warp assembles one image.  FUSED = false: the stand-alone kernel -- the image's connection tables are fetched
// from global memory up front.  FUSED = true: the match+assemble kernel -- matcher warps of the same CTA write each
// limb's rows straight into the shared-memory tables and raise s_ready[k]; the assembler acquires limb k's flag right
// before it consumes the limb, so matching limbs k+1.. overlaps assembling limb k.
template <bool FUSED>
__device__ __forceinline__ void assemble_image(const AssembleArgs &a, unsigned char *smem_raw, uint64_t &bar, int n, int img_in_call,
                                               int lane, const int *s_ready) {
    const Workspace &ws = a.ws;
    const int K = ws.K, L = ws.L, capP = ws.capP, capR = ws.capR;

    // shared memory: [conn_score | conn_norm | conn_ij | conn_count] of this image, then the person table
    const size_t LC = (size_t)L * capP;
    double *s_cs = reinterpret_cast<double *>(smem_raw);
    double *s_cn = s_cs + LC;
    uint32_t *s_cij = reinterpret_cast<uint32_t *>(s_cn + LC);
    int *s_cc = reinterpret_cast<int *>(s_cij + LC);
    unsigned char *table_base = smem_raw + assemble_conn_bytes(L, capP);

    PersonTable t = make_person_table(table_base, K, capP, capR);
    if (FUSED) {  // the CTA staged the coordinates behind the person table (match_assemble_kernel)
        t.px = reinterpret_cast<const double *>(table_base + assemble_smem_bytes(K, capP, capR));
        t.py = t.px + (size_t)K * capP;
    } else {
        t.px = ws.peak_x + (size_t)n * K * capP;
        t.py = ws.peak_y + (size_t)n * K * capP;
    }

    // Everything this image needs from global memory is fetched up front -- the connection tables by the bulk-copy
    // engine -- so that the serial limb loop below never waits on L2.
    const size_t img_conn = (size_t)n * LC;
    if (FUSED) {
        // the matcher warps fill the tables
    } else if (a.use_bulk) {
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
    if (!FUSED)
        for (int k = lane; k < L; k += 32) s_cc[k] = ws.conn_count[(size_t)n * L + k];
    if (!FUSED) init_person_table(t, ws, n, lane, 32, nullptr);  // fused: done by the whole CTA before the roles split
    __syncwarp();
    if (!FUSED && a.use_bulk) mbar_wait(&bar, 0);

    int nrows = 0;
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
        const int A = ws.limbs[2 * k], B = ws.limbs[2 * k + 1];
        for (int chunk = 0; chunk < cc && !overflow; chunk += 32) {
            const int mine_c = min(chunk + lane, cc - 1);
            const uint32_t my_ij = s_cij[k * capP + mine_c];
            const double my_s = s_cs[k * capP + mine_c], my_len = s_cn[k * capP + mine_c];
            const int in_chunk = min(32, cc - chunk);
            const int ia = (int)(my_ij >> 16), jb = (int)(my_ij & 0xffff);
            const int birth = (k << 8) | (chunk + lane);
            uint32_t pending = in_chunk == 32 ? 0xffffffffu : ((1u << in_chunk) - 1u);
            while (pending) {
                const bool mine = (pending >> lane) & 1u;
                int ra = -1, rb = -1;
                if (mine) {
                    ra = t.owner[A * capP + ia];
                    rb = t.owner[B * capP + jb];
                    if (ra >= 0) atomicMin(&t.touch[ra], lane);
                    if (rb >= 0 && rb != ra) atomicMin(&t.touch[rb], lane);
                }
                __syncwarp();
                // lowest pending toucher of every row it matches (the lowest pending lane always qualifies)
                const bool eligible = mine && (ra < 0 || t.touch[ra] == lane) && (rb < 0 || t.touch[rb] == lane);
                const uint32_t emask = __ballot_sync(0xffffffffu, eligible);
                if (mine) {
                    if (ra >= 0) t.touch[ra] = 0x7fffffff;
                    if (rb >= 0) t.touch[rb] = 0x7fffffff;
                }
                const uint32_t creates = __ballot_sync(0xffffffffu, eligible && ra < 0 && rb < 0);
                if (nrows + __popc(creates) > capR) {
                    flags |= kStRowOverflow;
                    overflow = true;
                    break;
                }
                __syncwarp();
                if (eligible)
                    flags |= apply_connection(t, a, A, B, ia, jb, my_s, my_len, nrows + __popc(creates & ((1u << lane) - 1u)), birth);
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
    if (FUSED && overflow) {  // the matchers may still be writing the tables this warp is about to reuse as staging space
        for (int k = 0; k < L; k++) {
            int r;
            do {
                asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(r) : "r"(smem_u32(s_ready + k)) : "memory");
                if (!r) __nanosleep(20);
            } while (!r);
        }
    }
    flags = __reduce_or_sync(0xffffffffu, flags);

    // ---- prune (:491-496) + outputs.  Kept rows keep their relative order.
    const int RS = K + 2, J = ws.J;
    double *g_subset = ws.subset + (size_t)n * capR * RS * 2;
    double *g_xy = ws.people_xy + (size_t)n * capR * J * 2;
    double *g_score = ws.people_score + (size_t)n * capR;
    const double *g_px = t.px, *g_py = t.py;
    // wire record (include/spgroup.h): rows are staged in shared memory -- the connection tables are dead by now -- and
    // leave in one coalesced copy, so a record in a peer GPU's memory costs a few 128-byte NVLink writes per image
    const int WR = 2 * J + 2;  // x,y per joint, person score, presence mask
    const bool wire_on = ws.wire != nullptr && (size_t)min(ws.wire_rows, capR) * WR * sizeof(double) <= assemble_conn_bytes(L, capP);
    double *s_wire = reinterpret_cast<double *>(smem_raw);
    // keep flags first (reusing `touch`), then each kept row's output position = number of kept rows born earlier
    for (int j = lane; j < nrows; j += 32) {
        bool keep = false;
        if (t.alive[j]) {
            const int cnt = t.cnt[j];
            keep = !(cnt < a.min_parts || __ddiv_rn(t.total[j], (double)cnt) < a.min_mean_score);
        }
        t.touch[j] = keep ? 1 : 0;
    }
    __syncwarp();
    int out = 0;
    for (int j = 0; j < nrows; j++) out += t.touch[j];
    for (int j = lane; j < nrows; j += 32) {
        if (!t.touch[j]) continue;
        const int mine = t.birth[j];
        int o = 0;
        for (int u = 0; u < nrows; u++) o += (t.touch[u] && t.birth[u] < mine);
        double *row = g_subset + (size_t)o * RS * 2;
        const uint32_t mj = t.mask[j];
        for (int c = 0; c < K; c++) {
            const bool has = (mj >> c) & 1u;
            row[c * 2 + 0] = has ? (double)t.id[c * capR + j] : -1.0;
            row[c * 2 + 1] = has ? t.sc[c * capR + j] : -1.0;
        }
        const double total = t.total[j];
        row[K * 2 + 0] = total;
        row[K * 2 + 1] = -1.0;
        row[(K + 1) * 2 + 0] = (double)t.cnt[j];
        row[(K + 1) * 2 + 1] = t.maxlen[j];
        const double pscore = __dsub_rn(1.0, __ddiv_rn(1.0, total));  // :541
        g_score[o] = pscore;
        const bool wrow = wire_on && o < ws.wire_rows;
        unsigned long long present = 0ull;
        if (wrow) s_wire[(size_t)o * WR + 2 * J] = pscore;
        for (int g = 0; g < J; g++) {                         // :523-539
            const int part = ws.out_from_part[g];
            const int id = ((mj >> part) & 1u) ? t.id[part * capR + j] : -1;
            double x = 0.0, y = 0.0;
            if (id >= 0) {
                const int idx = id - t.off[part];
                x = g_px[part * capP + idx];
                y = g_py[part * capP + idx];
                present |= 1ull << g;
            }
            g_xy[((size_t)o * J + g) * 2 + 0] = x;
            g_xy[((size_t)o * J + g) * 2 + 1] = y;
            if (wrow) {
                s_wire[(size_t)o * WR + 2 * g + 0] = x;
                s_wire[(size_t)o * WR + 2 * g + 1] = y;
            }
        }
        if (wrow) reinterpret_cast<unsigned long long *>(s_wire)[(size_t)o * WR + 2 * J + 1] = present;
    }
    if (ws.wire != nullptr && (!wire_on || out > ws.wire_rows)) flags |= kStWireOverflow;
    uint32_t st_word = 0;
    if (lane == 0) {
        ws.n_persons[n] = out;
        st_word = flags ? (atomicOr(&ws.status[n], flags) | flags) : ws.status[n];
    }
    if (ws.wire != nullptr) {
        __syncwarp();
        const size_t rec_bytes = 8 + (size_t)ws.wire_rows * WR * sizeof(double);
        unsigned char *rec = ws.wire + (size_t)(ws.wire_first + (long long)img_in_call) * rec_bytes;
        const int wn = wire_on ? min(out, ws.wire_rows) : 0;
        if (lane == 0) {
            reinterpret_cast<int *>(rec)[0] = wn;
            reinterpret_cast<uint32_t *>(rec)[1] = st_word;
        }
        double *rows = reinterpret_cast<double *>(rec + 8);
        for (int i = lane; i < wn * WR; i += 32) rows[i] = s_wire[i];
    }
    SPG_TR(301, 0);
    if (a.wire_flag != nullptr) {  // last CTA done publishes the step (the threadFenceReduction pattern, system scope)
        __syncwarp();              // every lane's record stores are ordered before lane 0's fence
        if (lane == 0) {
            __threadfence_system();
            const unsigned int prev = atomicAdd(a.done_counter, 1u);
            if (prev == (unsigned int)a.n_images - 1u) {
                *a.done_counter = 0u;  // re-armed for the next launch (stream order separates the launches)
                __threadfence_system();
                asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.wire_flag), "l"(a.wire_flag_value) : "memory");
            }
        }
    }
}

__global__ void __launch_bounds__(kAssembleThreads) assemble_kernel(AssembleArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    if ((int)blockIdx.x >= a.n_images) return;
    assemble_image<false>(a, smem_raw, bar, a.image_base + blockIdx.x, blockIdx.x, threadIdx.x, nullptr);
}

}  // namespace spg
