// limb_match.cuh -- K2b: greedy per-limb-type bipartite assignment.
//
// Replaces the matching half of find_connections (/root/reference/evaluate.py:259-274): a stable
// descending sort by priority followed by a sequential scan that accepts a candidate iff neither of its
// end points is used, stopping at min(nA, nB) connections.
//
// Sequential greedy over a strict total order is the same as repeatedly taking the best remaining
// candidate whose end points are both free.  One WARP per (image, limb) does exactly that: each lane
// holds a strided slice of the survivors in registers, every round is a warp arg-max on the key
// (priority desc, i*nB + j asc) -- the reference's stable-sort order, ties keep (i-major, j-minor)
// generation order -- and the winner's end points are struck from two 128-bit masks.  Rows come out in
// acceptance order, which find_people depends on.  No shared memory, no block barrier.
#pragma once

#include "common.cuh"

namespace spg {

struct MatchArgs {
    int n_images, image_base;
    Workspace ws;
};

constexpr int kMatchThreads = 128;
constexpr int kMatchRegCands = 8;  // survivors cached per lane (x32 lanes); beyond that we re-read L2

struct MatchKey {
    double prio;
    int p;    // i*nB + j, generation order
    int idx;  // position in the candidate list, -1 = none
};

__device__ __forceinline__ bool key_better(double pa, int ia, double pb, int ib) {
    return pa > pb || (pa == pb && ia < ib);
}

__global__ void __launch_bounds__(kMatchThreads) limb_match_kernel(MatchArgs a) {
    const Workspace &ws = a.ws;
    const int lane = threadIdx.x & 31;
    const int w = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    if (w >= a.n_images * ws.L) return;
    const int k = w % ws.L;
    const int n = a.image_base + w / ws.L;
    const size_t slot = (size_t)n * ws.L + k;
    const int nC = ws.cand_count[slot];
    if (nC < 0) {  // special_k
        if (lane == 0) ws.conn_count[slot] = -1;
        return;
    }
    const int pa = ws.limbs[2 * k], pb = ws.limbs[2 * k + 1];
    const int nA = min(ws.peak_count[(size_t)n * ws.K + pa], ws.capP);
    const int nB = min(ws.peak_count[(size_t)n * ws.K + pb], ws.capP);
    const int lim = min(nA, nB);
    const size_t cbase = slot * ws.capC;
    const size_t obase = slot * ws.capP;
    const size_t baseA = ((size_t)n * ws.K + pa) * ws.capP, baseB = ((size_t)n * ws.K + pb) * ws.capP;

    // register cache of this lane's survivors: candidate c = lane + 32*r
    double r_prio[kMatchRegCands];
    uint32_t r_ij[kMatchRegCands];
#pragma unroll
    for (int r = 0; r < kMatchRegCands; r++) {
        const int cidx = lane + 32 * r;
        const bool ok = cidx < nC;
        r_prio[r] = ok ? ws.cand_prio[cbase + cidx] : 0.0;
        r_ij[r] = ok ? ws.cand_ij[cbase + cidx] : 0xffffffffu;  // 0xffffffff = dead
    }

    unsigned long long uA0 = 0, uA1 = 0, uB0 = 0, uB1 = 0;
    auto used = [&](uint32_t ij) {
        const int i = ij >> 16, j = ij & 0xffff;
        const unsigned long long ma = (i < 64 ? uA0 : uA1) >> (i & 63);
        const unsigned long long mb = (j < 64 ? uB0 : uB1) >> (j & 63);
        return ((ma | mb) & 1ull) != 0;
    };

    int m = 0;
    while (m < lim) {
        double bp = 0.0;
        int bi = 0x7fffffff, bidx = -1;
#pragma unroll
        for (int r = 0; r < kMatchRegCands; r++) {
            const uint32_t ij = r_ij[r];
            if (ij == 0xffffffffu) continue;
            if (used(ij)) {
                r_ij[r] = 0xffffffffu;
                continue;
            }
            const int p = (int)(ij >> 16) * nB + (int)(ij & 0xffff);
            if (bidx < 0 || key_better(r_prio[r], p, bp, bi)) {
                bp = r_prio[r];
                bi = p;
                bidx = lane + 32 * r;
            }
        }
        for (int cidx = lane + 32 * kMatchRegCands; cidx < nC; cidx += 32) {  // rare: > 256 survivors
            const uint32_t ij = ws.cand_ij[cbase + cidx];
            if (used(ij)) continue;
            const double pr = ws.cand_prio[cbase + cidx];
            const int p = (int)(ij >> 16) * nB + (int)(ij & 0xffff);
            if (bidx < 0 || key_better(pr, p, bp, bi)) {
                bp = pr;
                bi = p;
                bidx = cidx;
            }
        }
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            const double op = __shfl_xor_sync(0xffffffffu, bp, s);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, s);
            const int oidx = __shfl_xor_sync(0xffffffffu, bidx, s);
            if (oidx >= 0 && (bidx < 0 || key_better(op, oi, bp, bi))) {
                bp = op;
                bi = oi;
                bidx = oidx;
            }
        }
        if (bidx < 0) break;  // no candidate with both end points free
        const int i = bi / nB, j = bi - i * nB;
        if (i < 64) uA0 |= 1ull << i; else uA1 |= 1ull << (i - 64);
        if (j < 64) uB0 |= 1ull << j; else uB1 |= 1ull << (j - 64);
        if (lane == 0) {  // row [idA, idB, score, i, j, norm] (evaluate.py:267)
            const double vx = __dsub_rn(ws.peak_x[baseB + j], ws.peak_x[baseA + i]);
            const double vy = __dsub_rn(ws.peak_y[baseB + j], ws.peak_y[baseA + i]);
            ws.conn_ij[obase + m] = ((uint32_t)i << 16) | (uint32_t)j;
            ws.conn_score[obase + m] = ws.cand_score[cbase + bidx];
            ws.conn_norm[obase + m] = __dsqrt_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)));
        }
        m++;
    }
    if (lane == 0) ws.conn_count[slot] = m;
}

}  // namespace spg
