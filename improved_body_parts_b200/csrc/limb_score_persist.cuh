// limb_score_persist.cuh -- K2a, persistent warp-specialised form (f32 planes that fit a 3-deep ring).
//
// Same arithmetic and outputs as limb_score_kernel (limb_score.cuh: conservative f32 screen, then the
// reference's exact evaluation of the survivors, evaluate.py:211-255); different schedule.  The one-CTA-per-
// (image, limb) kernel pays a prologue per item, serialises load -> screen -> exact inside a CTA and leaves
// issue slots idle at every barrier.  Here one CTA per SM stays resident and walks over its items
// (item = image * L + limb, strided by the grid) through a ring of 3 plane slots with three roles:
//
//   loader   (warp 0)       waits for a free slot, issues the plane's bulk copy (TMA, SASS UBLKCP) onto the slot's
//                           `full` mbarrier, then publishes the item's two end-point lists, which it fetched into
//                           registers one item earlier (every slot of the capacity is fetched, so the loads do not
//                           wait for the counters).  Neither global latency sits on the critical path.
//   workers  (warps 1-31)   per item j: phase A (screen) of their share of the pairs -- survivors are appended,
//                           warp-aggregated, to one of two lists -- then they pick up phase B (exact) work of item
//                           j-1 in chunks of 32 survivors from a shared counter, so whichever warps are ahead do
//                           the exact evaluation while the others are already screening the next item.  The last
//                           worker through publishes the item's counters and recycles the list.
//
// All hand-offs are mbarriers (no __syncthreads after start-up); the per-m tables are built once per CTA.
#pragma once

#include "limb_score.cuh"

namespace spg {

constexpr int kPersistThreads = 1024;
constexpr int kPersistSlots = 3;
constexpr int kWorkerWarps = kPersistThreads / 32 - 1;  // 31
constexpr int kPersistMaxCapP = 64;
constexpr int kPersistListCap = 2048;  // survivors queued per item; the (rare) excess is evaluated inline by the screener

struct PersistHdr {
    int nA, nB, npairs, n, k, special;
    uint32_t magic;
    int pad;
};

__host__ __device__ inline size_t persist_peaks_bytes(int capP) {
    return (((size_t)capP * (4 * sizeof(double) + 6 * sizeof(float) + 2)) + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t persist_tables_bytes() {
    return (((size_t)(kScreenMaxMid + 1) * (sizeof(double) + (kScreenSamples + 1) * sizeof(float) + 2)) + 15) & ~(size_t)15;
}
inline size_t persist_smem_bytes(size_t plane_bytes, int capP) {
    const size_t plane = (plane_bytes + 127) & ~(size_t)127;
    return kPersistSlots * plane + kPersistSlots * persist_peaks_bytes(capP) + persist_tables_bytes() +
           kPersistSlots * (size_t)kPersistListCap * sizeof(uint16_t) + 256;
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct PeakSlot {
    double *ax, *ay, *bx, *by;
    float *as, *bs, *fax, *fay, *fbx, *fby;
    unsigned char *ain, *bin;
};
__device__ __forceinline__ PeakSlot peak_slot(unsigned char *base, int capP) {
    PeakSlot p;
    p.ax = reinterpret_cast<double *>(base);
    p.ay = p.ax + capP;
    p.bx = p.ay + capP;
    p.by = p.bx + capP;
    p.as = reinterpret_cast<float *>(p.by + capP);
    p.bs = p.as + capP;
    p.fax = p.bs + capP;
    p.fay = p.fax + capP;
    p.fbx = p.fay + capP;
    p.fby = p.fbx + capP;
    p.ain = reinterpret_cast<unsigned char *>(p.fby + capP);
    p.bin = p.ain + capP;
    return p;
}

__global__ void __launch_bounds__(kPersistThreads, 1) limb_score_persist_kernel(ScoreArgs a, int n_items) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar_full[kPersistSlots], bar_free[kPersistSlots], bar_adone[kPersistSlots];
    __shared__ PersistHdr s_hdr[kPersistSlots];
    __shared__ int s_nsurv[kPersistSlots], s_ncand[kPersistSlots], s_done[kPersistSlots], s_bnext[kPersistSlots];
    __shared__ uint32_t s_flags[kPersistSlots];

    using T = float;
    const Workspace &ws = a.ws;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = a.H, W = a.W, capP = ws.capP, L = ws.L;
    const size_t plane_bytes = (size_t)H * W * sizeof(T);
    const size_t plane_stride = (plane_bytes + 127) & ~(size_t)127;
    unsigned char *peaks_base = smem_raw + kPersistSlots * plane_stride;
    const size_t peaks_stride = persist_peaks_bytes(capP);
    unsigned char *tables = peaks_base + kPersistSlots * peaks_stride;
    double *s_rcp = reinterpret_cast<double *>(tables);
    float *s_ts = reinterpret_cast<float *>(s_rcp + (kScreenMaxMid + 1));
    float *s_inv64 = s_ts + (size_t)(kScreenMaxMid + 1) * kScreenSamples;
    signed char *s_maxfail = reinterpret_cast<signed char *>(s_inv64 + (kScreenMaxMid + 1));
    unsigned char *s_qn = reinterpret_cast<unsigned char *>(s_maxfail + (kScreenMaxMid + 1));
    uint16_t *s_list = reinterpret_cast<uint16_t *>(tables + persist_tables_bytes());  // [slots][kPersistListCap]
    const int list_stride = kPersistListCap;

    // ---- one-time set-up
    if (tid == 0) {
        for (int s = 0; s < kPersistSlots; s++) {
            mbar_init(&bar_full[s], 2);  // plane copy (expect_tx) + end-point lists
            mbar_init(&bar_free[s], kWorkerWarps);
            mbar_init(&bar_adone[s], kWorkerWarps);
            s_nsurv[s] = 0; s_ncand[s] = 0; s_done[s] = 0; s_flags[s] = 0; s_bnext[s] = 0;
        }
        fence_mbar_init();
    }
    if (tid >= 32 && tid < 32 + kScreenMaxMid + 1) {
        const int m = tid - 32;
        const double need = __dmul_rn(a.connect_ration, (double)m);  // :246 compares in f64
        int need_i = (int)need;
        if ((double)need_i < need) need_i++;
        s_maxfail[m] = (signed char)max(min(m - need_i, 127), -1);
        s_rcp[m] = m > 0 ? __ddiv_rn(1.0, (double)m) : 0.0;
        s_inv64[m] = m > 1 ? 1.0f / (float)(m - 1) : 0.0f;
        const int lo = m / 8, hi = m - 1 - lo;
        const int qn = max(0, min(kScreenSamples, hi - lo + 1));
        s_qn[m] = (unsigned char)qn;
        for (int q = 0; q < kScreenSamples; q++)
            s_ts[m * kScreenSamples + q] = (float)(qn > 1 ? lo + (q * (hi - lo)) / (qn - 1) : lo);
    }
    __syncthreads();

    const int G = gridDim.x;
    const int nj = ((int)blockIdx.x < n_items) ? (n_items - 1 - (int)blockIdx.x) / G + 1 : 0;
    const bool screen = a.screen && a.mid_num <= kScreenMaxMid && H <= kScreenMaxDim && W <= kScreenMaxDim;
    const T thre2 = (T)a.thre2;

    if (warp == 0) {
        // =========================== loader ===========================
        // Software-pipelined: the end-point lists of item j+1 are already in registers while item j is published, and
        // the plane copy is issued before the lists are written, so neither global latency sits on the critical path.
        constexpr int kMaxE = kPersistMaxCapP / 32;  // list entries per lane
        double r_xa[kMaxE], r_ya[kMaxE], r_xb[kMaxE], r_yb[kMaxE];
        float r_sa[kMaxE], r_sb[kMaxE];
        int r_cntA = 0, r_cntB = 0;
        auto fetch = [&](int j) {
            const int item = (int)blockIdx.x + j * G;
            const int n_local = item / L, k = item - n_local * L;
            const int n = a.image_base + n_local;
            const int pa = ws.limbs[2 * k], pb = ws.limbs[2 * k + 1];
            r_cntA = ws.peak_count[(size_t)n * ws.K + pa];
            r_cntB = ws.peak_count[(size_t)n * ws.K + pb];
            const size_t baseA = ((size_t)n * ws.K + pa) * capP, baseB = ((size_t)n * ws.K + pb) * capP;
#pragma unroll
            for (int u = 0; u < kMaxE; u++) {
                const int e = lane + 32 * u;
                if (e < capP) {  // whole capacity: independent of the counters
                    r_xa[u] = ws.peak_x[baseA + e]; r_ya[u] = ws.peak_y[baseA + e];
                    r_xb[u] = ws.peak_x[baseB + e]; r_yb[u] = ws.peak_y[baseB + e];
                    r_sa[u] = ws.peak_score[baseA + e]; r_sb[u] = ws.peak_score[baseB + e];
                }
            }
        };
        if (nj > 0) fetch(0);
        for (int j = 0; j < nj; j++) {
            const int s = j % kPersistSlots;
            if (j >= kPersistSlots) mbar_wait_sleep(&bar_free[s], ((j / kPersistSlots) - 1) & 1);
            const int item = (int)blockIdx.x + j * G;
            const int n_local = item / L, k = item - n_local * L;
            const int n = a.image_base + n_local;
            if (lane == 0) {  // plane first (arrival 1 of 2 on `full`, carries the byte count)
                const unsigned char *gplane = reinterpret_cast<const unsigned char *>(
                    reinterpret_cast<const T *>(a.paf) + (int64_t)n_local * a.img_stride + (int64_t)k * a.chan_stride);
                unsigned char *dst = smem_raw + s * plane_stride;
                mbar_expect_tx(&bar_full[s], (uint32_t)plane_bytes);
                for (size_t off = 0; off < plane_bytes; off += kBulkChunkBytes) {
                    const uint32_t bytes = (uint32_t)min((size_t)kBulkChunkBytes, plane_bytes - off);
                    bulk_g2s(dst + off, gplane + off, bytes, &bar_full[s]);
                }
            }
            PeakSlot ps = peak_slot(peaks_base + s * peaks_stride, capP);
#pragma unroll
            for (int u = 0; u < kMaxE; u++) {
                const int e = lane + 32 * u;
                if (e < capP) {
                    const double xa = r_xa[u], ya = r_ya[u], xb = r_xb[u], yb = r_yb[u];
                    ps.ax[e] = xa; ps.ay[e] = ya; ps.bx[e] = xb; ps.by[e] = yb;
                    ps.as[e] = r_sa[u]; ps.bs[e] = r_sb[u];
                    ps.fax[e] = (float)(xa * 64.0); ps.fay[e] = (float)(ya * 64.0);
                    ps.fbx[e] = (float)(xb * 64.0); ps.fby[e] = (float)(yb * 64.0);
                    ps.ain[e] = xa >= 1.0 && xa <= (double)(W - 2) && ya >= 1.0 && ya <= (double)(H - 2);
                    ps.bin[e] = xb >= 1.0 && xb <= (double)(W - 2) && yb >= 1.0 && yb <= (double)(H - 2);
                }
            }
            const int nA = min(r_cntA, capP), nB = min(r_cntB, capP);
            const bool special = nA == 0 || nB == 0;
            if (lane == 0) {
                PersistHdr h;
                h.nA = nA; h.nB = nB; h.npairs = (special || a.debug == 1) ? 0 : nA * nB; h.n = n; h.k = k; h.special = special;
                h.magic = nB > 1 ? 0xffffffffu / (uint32_t)nB + 1u : 0u;
                h.pad = 0;
                s_hdr[s] = h;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_full[s]);  // arrival 2 of 2: lists + header are in place
            if (j + 1 < nj) fetch(j + 1);              // in flight while the next iteration waits for its slot
        }
    } else {
        // =========================== workers ===========================
        const int tidA = tid - 32;
        constexpr int kStrideA = kWorkerWarps * 32;

        // phase B of item jb: exact evaluation of survivor chunks taken from a shared counter; every worker passes
        // through here exactly once per item and arrives on b_done / slot_free when it has finished what it took
        // exact evaluation of one pair of the item in slot s + candidate append (phase B; also the inline path of a
        // screener whose survivor does not fit the list)
        auto exact_one = [&](int s, const PersistHdr &h, const PeakSlot &ps, const T *plane, int p) {
            const int nB = h.nB;
            const int i = nB > 1 ? (int)__umulhi((uint32_t)p, h.magic) : p;
            const int jj = p - i * nB;
            PairGeom g{ps.ax, ps.ay, ps.bx, ps.by, ps.as, ps.bs, s_rcp};
            double score, prio;
            bool bad = false;
            const bool ok = score_pair_exact<T>(plane, H, W, a, g, i, jj, ps.ain[i] && ps.bin[jj], thre2, score, prio, bad);
            if (bad) atomicOr(&s_flags[s], kStSampleIndex);
            if (ok) {
                const size_t out_base = ((size_t)h.n * L + h.k) * ws.capC;
                const int pos = atomicAdd(&s_ncand[s], 1);
                if (pos < ws.capC) {
                    const uint32_t ij = ((uint32_t)i << 16) | (uint32_t)jj;
                    ws.cand_prio[out_base + pos] = prio;
                    ws.cand_score[out_base + pos] = score;
                    ws.cand_ij[out_base + pos] = ij;
                    const uint32_t b = __float_as_uint((float)prio);
                    const uint32_t ord = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
                    ws.cand_key[out_base + pos] = ((unsigned long long)ord << 32) | (unsigned long long)(~ij);
                }
            }
        };

        auto phase_b = [&](int jb) {
            const int s = jb % kPersistSlots, q = s;  // list and counters live with the plane slot
            mbar_wait_sleep(&bar_adone[q], (jb / kPersistSlots) & 1);  // all workers have screened item jb: the list is complete
            const PersistHdr h = s_hdr[s];
            const PeakSlot ps = peak_slot(peaks_base + s * peaks_stride, capP);
            const T *plane = reinterpret_cast<const T *>(smem_raw + s * plane_stride);
            const uint16_t *list = s_list + q * list_stride;
            const int ns = a.debug == 2 ? 0 : min(s_nsurv[q], kPersistListCap);
            const size_t slot = (size_t)h.n * L + h.k;
            for (;;) {
                int c = 0;
                if (lane == 0) c = atomicAdd(&s_bnext[q], 1);
                c = __shfl_sync(0xffffffffu, c, 0);
                const int t = c * 32 + lane;
                if (c * 32 >= ns) break;
                if (t < ns) exact_one(s, h, ps, plane, list[t]);
                __syncwarp();
            }
            if (lane == 0) {
                __threadfence_block();  // this warp's appends before its "done" tick (and the other warps' after it)
                const bool last = atomicAdd(&s_done[q], 1) == kWorkerWarps - 1;
                __threadfence_block();
                if (last) {  // last worker through: publish + recycle
                    const int total = s_ncand[q];
                    ws.cand_count[slot] = h.special ? -1 : min(total, ws.capC);
                    if (ws.surv_count) ws.surv_count[slot] = s_nsurv[q];
                    uint32_t f = s_flags[q];
                    if (total > ws.capC) f |= kStCandOverflow;
                    if (f) atomicOr(&ws.status[h.n], f);
                    s_nsurv[q] = 0; s_ncand[q] = 0; s_done[q] = 0; s_flags[q] = 0; s_bnext[q] = 0;
                }
                mbar_arrive(&bar_free[s]);  // slot, list and counters may be reused once every worker has been here
            }
        };

        for (int j = 0; j < nj; j++) {
            const int s = j % kPersistSlots, q = s;
            // `full` also means the slot's list and counters are recycled: the loader refilled the slot only after every
            // worker had finished phase B of item j-3
            mbar_wait_sleep(&bar_full[s], (j / kPersistSlots) & 1);
            const PersistHdr h = s_hdr[s];
            const PeakSlot ps = peak_slot(peaks_base + s * peaks_stride, capP);
            const T *plane = reinterpret_cast<const T *>(smem_raw + s * plane_stride);
            uint16_t *list = s_list + q * list_stride;
            const int nB = h.nB;
            // ---- phase A (screen) of item j
            for (int base = 0; base + (tidA & ~31) < h.npairs; base += kStrideA) {  // warps without pairs skip the pass
                const int p = base + tidA;
                bool keep = false;
                if (p < h.npairs) {
                    keep = true;
                    if (screen) {
                        const int i = nB > 1 ? (int)__umulhi((uint32_t)p, h.magic) : p;
                        const int jj = p - i * nB;
                        if (ps.ain[i] && ps.bin[jj]) {
                            const float ax64 = ps.fax[i], ay64 = ps.fay[i];
                            const float dx64 = ps.fbx[jj] - ax64, dy64 = ps.fby[jj] - ay64;
                            const float n2 = (dx64 * dx64 + dy64 * dy64) * (1.0f / 4096.0f);  // px^2
                            if (n2 > 1e-6f) {
                                const float qf = n2 * rsqrtf(n2) + 1.0f;  // approximate norm + 1
                                // branch-free: lanes with long and short pairs must reach the sample loop together
                                const float r = rintf(qf);
                                const bool longp = qf >= (float)a.mid_num + 0.51f;
                                int m = longp ? a.mid_num : min((int)r, a.mid_num);
                                if (!longp && !(fabsf(qf - r) < 0.49f)) m = -1;  // m within 0.01 of a rounding tie -> survive
                                asm volatile("" : "+r"(m));  // keep ONE copy of the sample loop (no specialisation on m == mid_num)
                                if (m >= 1) {
                                    const int maxfail = s_maxfail[m];
                                    const int qn = s_qn[m];
                                    const float inv = s_inv64[m];
                                    const float sx64 = dx64 * inv, sy64 = dy64 * inv;
                                    // +33 folded into the start point: with u = pos + 33 (1/64 px), the pixel is u >> 6 for every
                                    // sample that is not within {31,32,33} (mod 64) of a rounding boundary, i.e. u & 63 > 2
                                    const float ax64o = ax64 + 33.0f, ay64o = ay64 + 33.0f;
                                    const float *ts = s_ts + m * kScreenSamples;
                                    int fails = 0;
                                    for (int q2 = 0; q2 < qn; q2++) {
                                        const float tf = ts[q2];
                                        const int xu = __float2int_rn(__fmaf_rn(tf, sx64, ax64o));
                                        const int yu = __float2int_rn(__fmaf_rn(tf, sy64, ay64o));
                                        const T v = plane[(yu >> 6) * W + (xu >> 6)];
                                        fails += (((unsigned)xu & 63u) > 2u) && (((unsigned)yu & 63u) > 2u) && !(v > thre2);
                                    }
                                    keep = fails <= maxfail;
                                }
                            }
                        }
                    }
                }
                const uint32_t km = __ballot_sync(0xffffffffu, keep);
                if (km) {
                    int at = 0;
                    if (lane == 0) at = atomicAdd(&s_nsurv[q], __popc(km));
                    at = __shfl_sync(0xffffffffu, at, 0);
                    if (keep) {
                        const int at_me = at + __popc(km & ((1u << lane) - 1u));
                        if (at_me < kPersistListCap) list[at_me] = (uint16_t)p;
                        else exact_one(s, h, ps, plane, p);  // list full: evaluate right here
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_adone[q]);
            // ---- phase B (exact) work of the previous item, for whichever warps get here while chunks are left
            if (j >= 1) phase_b(j - 1);
        }
        if (nj >= 1) phase_b(nj - 1);
    }
}

}  // namespace spg
