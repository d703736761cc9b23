// assemble.cuh -- K3: person assembly + final prune + COCO-ordered output.
//
// Replaces find_people (/root/reference/evaluate.py:279-498) and the tail of process() (:523-543).
//
// The algorithm is a state machine that consumes connections strictly in (limb, acceptance) order, so it
// is serial per image; throughput comes from the batch (one warp per image, all images concurrent).
// The `subset` table lives in shared memory in structure-of-arrays form -- id[part][row],
// score[part][row], total[row], count[row], maxlen[row] -- so that the per-connection scan "which rows
// hold idA in slot A or idB in slot B" is one conflict-free shared-memory read per lane plus a ballot.
// np.delete of a merged row (:424) becomes a tombstone bit: row order, which decides which two rows a
// connection matches (:304-318) and the output order, is unchanged by that.
// Connections of the current limb are prefetched 32 at a time (one per lane) and broadcast by shuffle,
// keeping global-memory latency off the serial chain.
#pragma once

#include "common.cuh"

namespace spg {

struct AssembleArgs {
    int n_images, image_base;
    double len_rate, connection_tole, min_mean_score;
    int remove_recon, min_parts;
    Workspace ws;
};

constexpr int kAssembleThreads = 32;

inline size_t assemble_smem_bytes(int K, int capP, int capR) {
    return (size_t)K * capR * sizeof(double)      // sc
           + 2 * (size_t)capR * sizeof(double)    // total, maxlen
           + (size_t)K * capR * sizeof(int)       // id
           + (size_t)capR * sizeof(int)           // cnt
           + (size_t)K * capP * sizeof(float)     // peak scores
           + (size_t)(K + 1) * sizeof(int);       // part offsets
}

__device__ __forceinline__ double shfl_f64(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

__global__ void __launch_bounds__(kAssembleThreads) assemble_kernel(AssembleArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const Workspace &ws = a.ws;
    const int lane = threadIdx.x;
    if ((int)blockIdx.x >= a.n_images) return;
    const int n = a.image_base + blockIdx.x;
    const int K = ws.K, L = ws.L, capP = ws.capP, capR = ws.capR;

    double *s_sc = reinterpret_cast<double *>(smem_raw);
    double *s_total = s_sc + (size_t)K * capR;
    double *s_maxlen = s_total + capR;
    int *s_id = reinterpret_cast<int *>(s_maxlen + capR);
    int *s_cnt = s_id + (size_t)K * capR;
    float *s_ps = reinterpret_cast<float *>(s_cnt + capR);
    int *s_off = reinterpret_cast<int *>(s_ps + (size_t)K * capP);

    if (lane == 0) {
        int acc = 0;
        for (int c = 0; c < K; c++) {
            s_off[c] = acc;
            acc += min(ws.peak_count[(size_t)n * K + c], capP);
        }
        s_off[K] = acc;
    }
    for (int t = lane; t < K * capP; t += 32) s_ps[t] = ws.peak_score[(size_t)n * K * capP + t];
    __syncwarp();

    uint32_t alive[4] = {0, 0, 0, 0};  // tombstone mask, uniform across lanes
    int nrows = 0;
    uint32_t flags = 0;
    bool overflow = false;

    for (int k = 0; k < L && !overflow; k++) {
        const size_t slot = (size_t)n * L + k;
        const int cc = ws.conn_count[slot];
        if (cc < 0) continue;  // special_k (:290)
        const int A = ws.limbs[2 * k], B = ws.limbs[2 * k + 1];
        const int offA = s_off[A], offB = s_off[B];
        for (int chunk = 0; chunk < cc && !overflow; chunk += 32) {
            const int mine = chunk + lane;
            uint32_t my_ij = 0;
            double my_s = 0.0, my_len = 0.0;
            if (mine < cc) {
                my_ij = ws.conn_ij[slot * capP + mine];
                my_s = ws.conn_score[slot * capP + mine];
                my_len = ws.conn_norm[slot * capP + mine];
            }
            const int in_chunk = min(32, cc - chunk);
            for (int r = 0; r < in_chunk; r++) {
                const uint32_t ij = __shfl_sync(0xffffffffu, my_ij, r);
                const double s = shfl_f64(my_s, r);
                const double len = shfl_f64(my_len, r);
                const int ia = (int)(ij >> 16), jb = (int)(ij & 0xffff);
                const int idA = offA + ia, idB = offB + jb;

                // ---- which (at most two, lowest-index) live rows hold idA in slot A or idB in slot B (:304-318)
                int found = 0, j1 = -1, j2 = -1;
#pragma unroll
                for (int w = 0; w < 4; w++) {
                    if (w * 32 < nrows && found < 2) {
                        const int j = w * 32 + lane;
                        bool hit = false;
                        if (j < nrows && ((alive[w] >> lane) & 1u))
                            hit = (s_id[A * capR + j] == idA) || (s_id[B * capR + j] == idB);
                        uint32_t mask = __ballot_sync(0xffffffffu, hit);
                        while (mask && found < 2) {
                            const int b = __ffs(mask) - 1;
                            mask &= mask - 1;
                            if (found == 0) j1 = w * 32 + b; else j2 = w * 32 + b;
                            found++;
                        }
                    }
                }

                if (found == 1) {  // :320-383 -- always slot B of the matched row
                    const int j = j1;
                    const int oldB = s_id[B * capR + j];
                    const double scB = s_sc[B * capR + j];
                    const double ml = s_maxlen[j];
                    const double reach = __dmul_rn(a.len_rate, ml);
                    const double add = __dadd_rn((double)s_ps[B * capP + jb], s);
                    if (oldB == -1 && reach > len) {  // assign (:323-342)
                        if (lane == 0) {
                            s_id[B * capR + j] = idB;
                            s_sc[B * capR + j] = s;
                            s_cnt[j] += 1;
                            s_total[j] = __dadd_rn(s_total[j], add);
                            s_maxlen[j] = len > ml ? len : ml;
                        }
                    } else if (oldB != idB) {
                        if (!(scB >= s) && !(reach <= len)) {  // replace (:346-363)
                            if (lane == 0) {
                                const int oldIdx = oldB >= 0 ? oldB - offB : 0;
                                const double sub = __dadd_rn((double)s_ps[B * capP + oldIdx], scB);
                                s_total[j] = __dadd_rn(__dsub_rn(s_total[j], sub), add);
                                s_id[B * capR + j] = idB;
                                s_sc[B * capR + j] = s;
                                s_maxlen[j] = len > ml ? len : ml;
                            }
                        }
                    } else if (scB <= s) {  // same B, refresh its score (:368-380)
                        if (lane == 0) {
                            const double sub = __dadd_rn((double)s_ps[B * capP + jb], scB);
                            s_total[j] = __dadd_rn(__dsub_rn(s_total[j], sub), add);
                            s_sc[B * capR + j] = s;
                            s_maxlen[j] = len > ml ? len : ml;
                        }
                    }
                    __syncwarp();
                } else if (found == 2) {  // :385-460
                    const int id1 = lane < K ? s_id[lane * capR + j1] : -1;
                    const int id2 = lane < K ? s_id[lane * capR + j2] : -1;
                    const double sc1 = lane < K ? s_sc[lane * capR + j1] : 0.0;
                    const double sc2 = lane < K ? s_sc[lane * capR + j2] : 0.0;
                    const uint32_t both = __ballot_sync(0xffffffffu, id1 >= 0 && id2 >= 0);
                    if (both == 0) {  // disjoint -> merge j2 into j1 (:403-424)
                        double m = fmin(id1 >= 0 ? sc1 : INFINITY, id2 >= 0 ? sc2 : INFINITY);
#pragma unroll
                        for (int sft = 16; sft > 0; sft >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, sft));
                        const double ml1 = s_maxlen[j1];
                        if (!(s < __dmul_rn(a.connection_tole, m) || __dmul_rn(a.len_rate, ml1) <= len)) {
                            if (lane < K) {  // the "+1" trick (:415): absent slots are -1 in both columns
                                s_id[lane * capR + j1] = id1 + id2 + 1;
                                s_sc[lane * capR + j1] = __dadd_rn(sc1, __dadd_rn(sc2, 1.0));
                            }
                            if (lane == 0) {
                                s_total[j1] = __dadd_rn(__dadd_rn(s_total[j1], s_total[j2]), s);  // :419, :421
                                s_cnt[j1] += s_cnt[j2];
                                s_maxlen[j1] = len > ml1 ? len : ml1;  // keeps j1's own longest limb (:422)
                            }
                            alive[j2 >> 5] &= ~(1u << (j2 & 31));  // np.delete(subset, j2) (:424)
                        }
                    } else {  // overlapping rows (:426-460): only remove_recon > 0 has side effects
                        const uint32_t a_in_1 = __ballot_sync(0xffffffffu, id1 == idA);
                        const int k1 = a_in_1 ? idA : idB, k2 = a_in_1 ? idB : idA;
                        const uint32_t m1 = __ballot_sync(0xffffffffu, id1 == k1);
                        const uint32_t m2 = __ballot_sync(0xffffffffu, id2 == k2);
                        if (__popc(m1) != 1 || __popc(m2) != 1 || m1 == m2) {
                            flags |= kStAssert;  // the reference would raise (:437-439)
                        } else {
                            const int c1 = __ffs(m1) - 1, c2 = __ffs(m2) - 1;
                            const double e1 = s_sc[c1 * capR + j1], e2 = s_sc[c2 * capR + j2];
                            if (!(s < e1 && s < e2) && a.remove_recon > 0) {
                                int small_j = j1, rc = c1;
                                if (e1 > e2) { small_j = j2; rc = c2; }
                                if (lane == 0) {
                                    const int rid = s_id[rc * capR + small_j];
                                    const double sub = __dadd_rn((double)s_ps[rc * capP + (rid - s_off[rc])],
                                                                 s_sc[rc * capR + small_j]);
                                    s_total[small_j] = __dsub_rn(s_total[small_j], sub);
                                    s_id[rc * capR + small_j] = -1;
                                    s_sc[rc * capR + small_j] = -1.0;
                                    s_cnt[small_j] -= 1;
                                }
                            }
                        }
                    }
                    __syncwarp();
                } else {  // new person (:473-488)
                    if (nrows >= capR) {
                        flags |= kStRowOverflow;
                        overflow = true;
                        break;
                    }
                    const int j = nrows;
                    if (lane < K) {
                        s_id[lane * capR + j] = lane == A ? idA : (lane == B ? idB : -1);
                        s_sc[lane * capR + j] = (lane == A || lane == B) ? s : -1.0;
                    }
                    if (lane == 0) {
                        s_cnt[j] = 2;
                        s_maxlen[j] = len;
                        // builtin sum() of the two end-point scores, then + s (:484)
                        s_total[j] = __dadd_rn(__dadd_rn(__dadd_rn(0.0, (double)s_ps[A * capP + ia]), (double)s_ps[B * capP + jb]), s);
                    }
                    alive[j >> 5] |= 1u << (j & 31);
                    nrows++;
                    __syncwarp();
                }
            }
        }
    }

    // ---- prune (:491-496) + outputs.  Kept rows keep their relative order.
    const int RS = K + 2, J = ws.J;
    double *g_subset = ws.subset + (size_t)n * capR * RS * 2;
    double *g_xy = ws.people_xy + (size_t)n * capR * J * 2;
    double *g_score = ws.people_score + (size_t)n * capR;
    const double *g_px = ws.peak_x + (size_t)n * K * capP;
    const double *g_py = ws.peak_y + (size_t)n * K * capP;
    int out = 0;
    for (int j = 0; j < nrows; j++) {
        if (!((alive[j >> 5] >> (j & 31)) & 1u)) continue;
        const int cnt = s_cnt[j];
        const double total = s_total[j];
        if (cnt < a.min_parts || __ddiv_rn(total, (double)cnt) < a.min_mean_score) continue;
        double *row = g_subset + (size_t)out * RS * 2;
        if (lane < K) {
            row[lane * 2 + 0] = (double)s_id[lane * capR + j];
            row[lane * 2 + 1] = s_sc[lane * capR + j];
        }
        if (lane == 0) {
            row[K * 2 + 0] = total;
            row[K * 2 + 1] = -1.0;
            row[(K + 1) * 2 + 0] = (double)cnt;
            row[(K + 1) * 2 + 1] = s_maxlen[j];
            g_score[out] = __dsub_rn(1.0, __ddiv_rn(1.0, total));  // :541
        }
        for (int g = lane; g < J; g += 32) {  // :523-539
            const int part = ws.out_from_part[g];
            const int id = s_id[part * capR + j];
            double x = 0.0, y = 0.0;
            if (id >= 0) {
                const int idx = id - s_off[part];
                x = g_px[part * capP + idx];
                y = g_py[part * capP + idx];
            }
            g_xy[((size_t)out * J + g) * 2 + 0] = x;
            g_xy[((size_t)out * J + g) * 2 + 1] = y;
        }
        out++;
    }
    if (lane == 0) {
        ws.n_persons[n] = out;
        if (flags) atomicOr(&ws.status[n], flags);
    }
}

}  // namespace spg
