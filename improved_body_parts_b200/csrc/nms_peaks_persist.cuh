// nms_peaks_persist.cuh -- K1, persistent warp-specialised form (planes that fit a 3-deep ring, W % 4 == 0).
//
// Same results as nms_peaks_kernel (nms_peaks.cuh); different schedule.  4608 short-lived CTAs per launch spend
// most of their life in launch / first-copy latency and in a one-warp refinement tail that keeps their shared
// memory pinned.  Here one CTA per SM stays resident and walks over its (image, part) planes:
//
//   loader    (warp 0)       waits for a free plane slot (ring of 3) and issues the plane's bulk copy (TMA, SASS
//                            UBLKCP) onto the slot's `full` mbarrier.
//   scanners  (warps 1-28)   per plane: pass 1 queues the float4 groups of their slice that reach thre1, pass 2 runs
//                            the 8-neighbour test on the queue and appends peaks to the plane's list (ring of 6);
//                            then they release the plane slot -- the plane is not needed any more.
//   finishers (warps 29-31)  take planes round-robin: rank the list by raster index (= np.nonzero order), refine each
//                            peak from L2 (the plane was just streamed), write the outputs, recycle the list.
#pragma once

#include "nms_peaks.cuh"

namespace spg {

constexpr int kNmsPThreads = 1024;
constexpr int kNmsPSlots = 3;
constexpr int kNmsPFinishers = 3;
constexpr int kNmsPLists = 2 * kNmsPFinishers;
constexpr int kNmsPScanners = kNmsPThreads / 32 - 1 - kNmsPFinishers;  // 28
constexpr int kNmsPMaxIter = 5;  // 32-lane passes over a scanner's slice: 3 planes must fit in shared memory, so a slice is <= 160 float4 groups

inline size_t nms_persist_smem_bytes(int H, int W, int capP) {
    const size_t plane = (((size_t)H * W * sizeof(float)) + 127) & ~(size_t)127;
    const size_t gpw = ((size_t)H * W / 4 + kNmsPScanners - 1) / kNmsPScanners;
    const size_t queues = ((gpw * kNmsPScanners * sizeof(uint16_t)) + 15) & ~(size_t)15;
    return kNmsPSlots * plane + kNmsPLists * (size_t)capP * sizeof(uint32_t) + queues + 64;
}

__device__ __forceinline__ void mbar_arrive_plain(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(kNmsPThreads, 1) nms_peaks_persist_kernel(NmsArgs a, int n_items) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar_full[kNmsPSlots], bar_free[kNmsPSlots], bar_ready[kNmsPLists], bar_lfree[kNmsPLists];
    __shared__ int s_cnt[kNmsPLists];

    const Workspace &ws = a.ws;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = a.H, W = a.W, K = ws.K, capP = ws.capP;
    const size_t plane_bytes = (size_t)H * W * sizeof(float);
    const size_t plane_stride = (plane_bytes + 127) & ~(size_t)127;
    uint32_t *s_lists = reinterpret_cast<uint32_t *>(smem_raw + kNmsPSlots * plane_stride);  // [lists][capP]
    uint16_t *s_queues = reinterpret_cast<uint16_t *>(s_lists + kNmsPLists * (size_t)capP);
    const int W4 = W >> 2, groups = H * W4;
    // g / W4 without a division: groups < 2^16 here (launch condition), so umulhi(g, ceil(2^32 / W4)) is exact
    const uint32_t w4_magic = W4 > 1 ? 0xffffffffu / (uint32_t)W4 + 1u : 0u;
    const int gpw = (groups + kNmsPScanners - 1) / kNmsPScanners;

    if (tid == 0) {
        for (int s = 0; s < kNmsPSlots; s++) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_free[s], kNmsPScanners);
        }
        for (int l = 0; l < kNmsPLists; l++) {
            mbar_init(&bar_ready[l], kNmsPScanners);
            mbar_init(&bar_lfree[l], 1);
            s_cnt[l] = 0;
        }
        fence_mbar_init();
    }
    __syncthreads();

    const int G = gridDim.x;
    const int nj = ((int)blockIdx.x < n_items) ? (n_items - 1 - (int)blockIdx.x) / G + 1 : 0;

    if (warp == 0) {
        // =========================== loader ===========================
        if (lane == 0) {
            for (int j = 0; j < nj; j++) {
                const int s = j % kNmsPSlots;
                if (j >= kNmsPSlots) mbar_wait_sleep(&bar_free[s], ((j / kNmsPSlots) - 1) & 1);
                const int item = (int)blockIdx.x + j * G;
                const int n_local = item / K, c = item - n_local * K;
                const unsigned char *src =
                    reinterpret_cast<const unsigned char *>(a.heat + (int64_t)n_local * a.img_stride + (int64_t)c * a.chan_stride);
                unsigned char *dst = smem_raw + s * plane_stride;
                mbar_expect_tx(&bar_full[s], (uint32_t)plane_bytes);
                for (size_t off = 0; off < plane_bytes; off += 32768) {
                    const uint32_t bytes = (uint32_t)min((size_t)32768, plane_bytes - off);
                    bulk_g2s(dst + off, src + off, bytes, &bar_full[s]);
                }
            }
        }
    } else if (warp <= kNmsPScanners) {
        // =========================== scanners ===========================
        const int w = warp - 1;
        uint16_t *wq = s_queues + (size_t)w * gpw;
        const int g_lo = w * gpw, g_hi = min(g_lo + gpw, groups);
        const float thr = a.thr;
        for (int j = 0; j < nj; j++) {
            const int s = j % kNmsPSlots, l = j % kNmsPLists;
            mbar_wait_sleep(&bar_full[s], (j / kNmsPSlots) & 1);
            if (j >= kNmsPLists) mbar_wait_sleep(&bar_lfree[l], ((j / kNmsPLists) - 1) & 1);
            const float *buf = reinterpret_cast<const float *>(smem_raw + s * plane_stride);
            uint32_t *list = s_lists + (size_t)l * capP;
            // ---- pass 1: queue the float4 groups of this warp's slice that reach thre1.  First only the votes (one
            // load, three max, one compare, one ballot per 128 elements -- the common case is an empty mask), then the
            // queue from the masks.
            uint32_t am[kNmsPMaxIter];
#pragma unroll
            for (int it = 0; it < kNmsPMaxIter; it++) {
                const int g = g_lo + it * 32 + lane;
                bool act = false;
                if (g < g_hi) {
                    const float4 c4 = *reinterpret_cast<const float4 *>(buf + 4 * (size_t)g);
                    act = fmaxf(fmaxf(c4.x, c4.y), fmaxf(c4.z, c4.w)) >= thr;
                }
                am[it] = __ballot_sync(0xffffffffu, act);
            }
            int nq = 0;
#pragma unroll
            for (int it = 0; it < kNmsPMaxIter; it++) {
                const uint32_t m = am[it];
                if (m) {  // warp-uniform
                    if ((m >> lane) & 1u) wq[nq + __popc(m & ((1u << lane) - 1u))] = (uint16_t)(g_lo + it * 32 + lane);
                    nq += __popc(m);
                }
            }
            __syncwarp();
            // ---- pass 2: 8-neighbour test (neighbours clamped to the image == window clipped to the image)
            for (int q = lane; q < nq; q += 32) {
                const int g = wq[q];
                const int y = W4 > 1 ? (int)__umulhi((uint32_t)g, w4_magic) : g, xq = g - y * W4;
                const int x0 = 4 * xq;
                const float *rc = buf + (size_t)y * W;
                const float *ru = buf + (size_t)max(y - 1, 0) * W;
                const float *rd = buf + (size_t)min(y + 1, H - 1) * W;
                const float4 c4 = *reinterpret_cast<const float4 *>(rc + x0);
                const float4 u4 = *reinterpret_cast<const float4 *>(ru + x0);
                const float4 d4 = *reinterpret_cast<const float4 *>(rd + x0);
                const int xl = max(x0 - 1, 0), xr = min(x0 + 4, W - 1);
                const float U[6] = {ru[xl], u4.x, u4.y, u4.z, u4.w, ru[xr]};
                const float C[6] = {rc[xl], c4.x, c4.y, c4.z, c4.w, rc[xr]};
                const float D[6] = {rd[xl], d4.x, d4.y, d4.z, d4.w, rd[xr]};
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const float v = C[e + 1];
                    // keep = (hmax == heat) & (heat >= thre) (util.py:182); np.nonzero(heat * keep) drops exact zeros
                    const bool pk = (v >= thr) & (v != 0.0f) & (U[e] <= v) & (U[e + 1] <= v) & (U[e + 2] <= v) &
                                    (C[e] <= v) & (C[e + 2] <= v) & (D[e] <= v) & (D[e + 1] <= v) & (D[e + 2] <= v);
                    if (pk) {
                        const int pos = atomicAdd(&s_cnt[l], 1);
                        if (pos < capP) list[pos] = (uint32_t)(y * W + x0 + e);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_plain(&bar_free[s]);   // the plane is not needed any more
                mbar_arrive_plain(&bar_ready[l]);  // this warp's peaks are in the list
            }
        }
    } else {
        // =========================== finishers ===========================
        const int f = warp - 1 - kNmsPScanners;
        const int R = a.radius;
        for (int j = f; j < nj; j += kNmsPFinishers) {
            const int l = j % kNmsPLists;
            mbar_wait_sleep(&bar_ready[l], (j / kNmsPLists) & 1);
            const uint32_t *list = s_lists + (size_t)l * capP;
            const int item = (int)blockIdx.x + j * G;
            const int n_local = item / K, c = item - n_local * K;
            const int n = a.image_base + n_local;
            const float *plane = a.heat + (int64_t)n_local * a.img_stride + (int64_t)c * a.chan_stride;  // L2-hot
            const int total = s_cnt[l];
            const int np = min(total, capP);
            const size_t out_base = ((size_t)n * K + c) * capP;
            for (int t = lane; t < np; t += 32) {
                const uint32_t mine = list[t];
                int rank = 0;  // raster index rank == np.nonzero order (evaluate.py:193); indices are unique
                for (int u = 0; u < np; u++) rank += list[u] < mine;
                const int lin = (int)mine;
                const int y = lin / W, x = lin - y * W;
                double rx, ry;
                float sc;
                uint32_t anchor = ((uint32_t)y << 16) | (uint32_t)x;
                if (y + R + 1 > H || y - R < 0 || x + R + 1 > W || x - R < 0) {
                    rx = (double)x;  // util.py:201-202: the box leaves the image -> integer anchor, raw map value
                    ry = (double)y;
                    sc = plane[(size_t)y * W + x];
                    anchor |= 0x80000000u;
                } else {
                    switch (R) {  // util.py:204-211
                        case 0: refine_box<0>(plane, W, x, y, rx, ry, sc); break;
                        case 1: refine_box<1>(plane, W, x, y, rx, ry, sc); break;
                        case 2: refine_box<2>(plane, W, x, y, rx, ry, sc); break;
                        case 3: refine_box<3>(plane, W, x, y, rx, ry, sc); break;
                        default: refine_box<4>(plane, W, x, y, rx, ry, sc); break;
                    }
                }
                ws.peak_x[out_base + rank] = rx;
                ws.peak_y[out_base + rank] = ry;
                ws.peak_score[out_base + rank] = sc;
                ws.peak_anchor[out_base + rank] = anchor;
            }
            __syncwarp();
            if (lane == 0) {
                ws.peak_count[(size_t)n * K + c] = total;
                if (total > capP) atomicOr(&ws.status[n], kStPeakOverflow);
                s_cnt[l] = 0;
                mbar_arrive_plain(&bar_lfree[l]);  // list + counter may be reused (plane j + kNmsPLists)
            }
        }
    }
}

}  // namespace spg
