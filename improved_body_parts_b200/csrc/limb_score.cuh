// limb_score.cuh -- K2a: line-integral scoring of every candidate keypoint pair of every limb.
//
// Replaces the scoring half of find_connections (/root/reference/evaluate.py:211-255).
//
// One CTA per (image, limb).  The limb's body-part plane (H*W elements, one contiguous span) is staged
// into shared memory by the bulk-copy engine (TMA, SASS UBLKCP) in a few large chunks on one mbarrier,
// while the threads stage the two end-point peak lists; HBM is read exactly once per plane element and
// all <= nA*nB*mid_num nearest-neighbour gathers hit shared memory.  Planes that do not fit in shared
// memory (e.g. 512x512) are sampled through L2 instead (STAGE = false).
//
// One THREAD per candidate pair, not one warp: the reference sums the <= mid_num samples of a pair
// sequentially (Python sum() over np.float32, evaluate.py:241), and a shuffle-tree reduction would round
// differently and can flip threshold / ordering decisions downstream.  The warp-level primitive used here
// is the aggregated append of surviving candidates.  Survivors are written unordered; the matcher orders
// them by (priority desc, i*nB+j asc), which is the reference's stable-sort order (evaluate.py:259).
#pragma once

#include "common.cuh"

namespace spg {

struct ScoreArgs {
    const void *paf;
    int64_t img_stride, chan_stride;  // elements
    int H, W, image_base, mid_num;
    double image_extent, thre2, connect_ration;
    Workspace ws;
};

constexpr int kScoreThreads = 256;
constexpr uint32_t kBulkChunkBytes = 32768;

inline size_t score_smem_bytes(size_t plane_bytes, int capP) {
    const size_t plane = (plane_bytes + 127) & ~(size_t)127;
    return plane + (size_t)capP * (4 * sizeof(double) + 2 * sizeof(float));
}

template <typename T, bool STAGE>
__global__ void __launch_bounds__(kScoreThreads) limb_score_kernel(ScoreArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ int s_count;
    __shared__ uint32_t s_flags;

    const Workspace &ws = a.ws;
    const int tid = threadIdx.x;
    const int k = blockIdx.x % ws.L;
    const int n_local = blockIdx.x / ws.L;
    const int n = a.image_base + n_local;
    const int H = a.H, W = a.W;
    const int pa = ws.limbs[2 * k], pb = ws.limbs[2 * k + 1];
    const int nA = min(ws.peak_count[(size_t)n * ws.K + pa], ws.capP);
    const int nB = min(ws.peak_count[(size_t)n * ws.K + pb], ws.capP);
    const size_t slot = (size_t)n * ws.L + k;
    if (nA == 0 || nB == 0) {  // special_k (evaluate.py:272-274)
        if (tid == 0) ws.cand_count[slot] = -1;
        return;
    }

    const T *gplane = reinterpret_cast<const T *>(a.paf) + (int64_t)n_local * a.img_stride + (int64_t)k * a.chan_stride;
    const size_t plane_bytes = (size_t)H * W * sizeof(T);
    T *splane = reinterpret_cast<T *>(smem_raw);
    unsigned char *after = smem_raw + (STAGE ? ((plane_bytes + 127) & ~(size_t)127) : 0);
    double *s_ax = reinterpret_cast<double *>(after);
    double *s_ay = s_ax + ws.capP;
    double *s_bx = s_ay + ws.capP;
    double *s_by = s_bx + ws.capP;
    float *s_as = reinterpret_cast<float *>(s_by + ws.capP);
    float *s_bs = s_as + ws.capP;

    if (tid == 0) {
        s_count = 0;
        s_flags = 0;
        if (STAGE) {
            mbar_init(&bar, 1);
            fence_mbar_init();
            mbar_expect_tx(&bar, (uint32_t)plane_bytes);
            for (size_t off = 0; off < plane_bytes; off += kBulkChunkBytes) {
                const uint32_t bytes = (uint32_t)min((size_t)kBulkChunkBytes, plane_bytes - off);
                bulk_g2s(smem_raw + off, reinterpret_cast<const unsigned char *>(gplane) + off, bytes, &bar);
            }
        }
    }
    // end-point lists (refined float coordinates + peak scores), overlapped with the plane copy
    const size_t baseA = ((size_t)n * ws.K + pa) * ws.capP, baseB = ((size_t)n * ws.K + pb) * ws.capP;
    for (int i = tid; i < nA; i += kScoreThreads) {
        s_ax[i] = ws.peak_x[baseA + i];
        s_ay[i] = ws.peak_y[baseA + i];
        s_as[i] = ws.peak_score[baseA + i];
    }
    for (int j = tid; j < nB; j += kScoreThreads) {
        s_bx[j] = ws.peak_x[baseB + j];
        s_by[j] = ws.peak_y[baseB + j];
        s_bs[j] = ws.peak_score[baseB + j];
    }
    __syncthreads();
    if (STAGE) mbar_wait(&bar, 0);
    const T *plane = STAGE ? splane : gplane;

    const T thre2 = (T)a.thre2;  // f32 plane: `> thre2` is an f32 compare against (float)thre2
    const size_t out_base = slot * ws.capC;
    const int npairs = nA * nB;
    for (int p = tid; p < npairs; p += kScoreThreads) {
        const int i = p / nB, j = p - i * nB;
        const double ax = s_ax[i], ay = s_ay[i], bx = s_bx[j], by = s_by[j];
        const double vx = __dsub_rn(bx, ax), vy = __dsub_rn(by, ay);                           // :224
        const double norm = __dsqrt_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)));       // :225
        if (norm == 0.0) continue;                                                             // :228-230
        int m = __double2int_rn(__dadd_rn(norm, 1.0));                                         // :226 round() half-even
        m = min(m, a.mid_num);
        // np.linspace(A, B, m): step = delta/(m-1); y_t = t*step + start (two roundings); y_{m-1} = stop
        const double stepx = m > 1 ? __ddiv_rn(vx, (double)(m - 1)) : 0.0;
        const double stepy = m > 1 ? __ddiv_rn(vy, (double)(m - 1)) : 0.0;
        T sum = (T)0;
        int above = 0;
        bool bad = false;
        for (int t = 0; t < m; t++) {
            double sx, sy;
            if (t == m - 1 && m > 1) {
                sx = bx;
                sy = by;
            } else {
                sx = __dadd_rn(__dmul_rn((double)t, stepx), ax);
                sy = __dadd_rn(__dmul_rn((double)t, stepy), ay);
            }
            int yi = __double2int_rn(sy), xi = __double2int_rn(sx);                            // :235 nearest neighbour
            if (yi < 0) yi += H;  // numpy index semantics: negatives wrap once
            if (xi < 0) xi += W;
            if (yi < 0 || yi >= H || xi < 0 || xi >= W) {  // the reference would raise IndexError
                bad = true;
                break;
            }
            const T v = plane[(size_t)yi * W + xi];
            sum = sum + v;  // sequential, in sample order, in the plane's precision (:241)
            above += v > thre2;
        }
        if (bad) {
            atomicOr(&s_flags, kStSampleIndex);
            continue;
        }
        // :241 -- `image_width` is the image HEIGHT at the call site (:510)
        const double prior = __dsub_rn(__ddiv_rn(__dmul_rn(0.5, a.image_extent), norm), 1.0);
        double score, prio;
        if (sizeof(T) == 4) {
            float s = __fdiv_rn((float)sum, (float)m);
            s = __fadd_rn(s, prior < 0.0 ? __double2float_rn(prior) : 0.0f);  // f32 + weak Python float
            const float pr = __fadd_rn(__fadd_rn(__fmul_rn(0.5f, s), __fmul_rn(0.25f, s_as[i])), __fmul_rn(0.25f, s_bs[j]));
            score = (double)s;
            prio = (double)pr;
        } else {
            score = __dadd_rn(__ddiv_rn((double)sum, (double)m), prior < 0.0 ? prior : 0.0);
            prio = __dadd_rn(__dadd_rn(__dmul_rn(0.5, score), (double)__fmul_rn(0.25f, s_as[i])),
                             (double)__fmul_rn(0.25f, s_bs[j]));
        }
        const bool crit1 = (double)above >= __dmul_rn(a.connect_ration, (double)m);            // :246
        const bool crit2 = score > 0.0;                                                        // :251
        if (crit1 && crit2) {
            const int pos = atomicAdd(&s_count, 1);  // warp-aggregated by ptxas (REDUX + one ATOMS)
            if (pos < ws.capC) {
                ws.cand_prio[out_base + pos] = prio;
                ws.cand_score[out_base + pos] = score;
                ws.cand_ij[out_base + pos] = ((uint32_t)i << 16) | (uint32_t)j;
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        const int total = s_count;
        ws.cand_count[slot] = min(total, ws.capC);
        uint32_t f = s_flags;
        if (total > ws.capC) f |= kStCandOverflow;
        if (f) atomicOr(&ws.status[n], f);
    }
}

}  // namespace spg
