// match_assemble.cuh -- K2b + K3 fused: greedy matching (evaluate.py:259-274) and person assembly
// (evaluate.py:279-498 + the process() tail :523-543 + the wire record) of one image in ONE CTA.
//
// Back to back, limb_match (7 680 independent one-warp chains, ~37 us) and assemble (256 one-warp chains of 30 limbs,
// ~54 us) are both latency-bound and use < 5 % of the machine; the second cannot start before the first has finished
// every limb of every image.  Inside one CTA per image the dependency is per limb: warps 1..7 match the image's
// limbs (limb k by warp 1 + k % 7, so limb k is ready long before the assembler needs it) and write each limb's
// rows into the assembler's SHARED-MEMORY tables; warp 0 assembles, acquiring a per-limb flag just before it consumes
// the limb.  The connection tables never make a round trip through L2 on the critical path (they are still stored to
// global memory for the stage-wise API), one launch and one dependent wave disappear, and the assembly of limb 0
// starts as soon as the first matcher is done.  Same device functions as the two stand-alone kernels: same results.
#pragma once

#include "assemble.cuh"
#include "limb_match.cuh"

namespace spg {

constexpr int kMAMatchWarps = 7;   // default; the launch may use 1..15 (blockDim.x = 32 * (1 + matchers))
constexpr int kMAThreads = 32 * (1 + kMAMatchWarps);
constexpr int kMAMaxThreads = 512;

inline size_t match_assemble_smem_bytes(int K, int L, int capP, int capR) {  // tables + person table + staged coordinates
    return assemble_conn_bytes(L, capP) + assemble_smem_bytes(K, capP, capR) + 2 * (size_t)K * capP * sizeof(double);
}

__global__ void __launch_bounds__(kMAMaxThreads) match_assemble_kernel(AssembleArgs a, int keys_valid) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;  // unused in the fused form (the matchers fill the tables)
    __shared__ int s_ready[kMaxLimbs];
    const Workspace &ws = a.ws;
    if ((int)blockIdx.x >= a.n_images) return;
    const int n = a.image_base + blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = ws.L, capP = ws.capP;
    SPG_TR(warp == 0 ? 0 : 2 + warp, 0);
    if (tid < L) s_ready[tid] = 0;
    {   // CTA-wide prologue: person table + the image's refined coordinates into shared memory
        unsigned char *table_base = smem_raw + assemble_conn_bytes(L, capP);
        const PersonTable t = make_person_table(table_base, ws.K, capP, ws.capR);
        init_person_table(t, ws, n, tid, (int)blockDim.x, reinterpret_cast<double *>(table_base + assemble_smem_bytes(ws.K, capP, ws.capR)));
    }
    __syncthreads();
    if (warp == 0) SPG_TR(1, 0);
    if (warp == 0) {
        assemble_image<true>(a, smem_raw, bar, n, blockIdx.x, lane, s_ready);
        return;
    }
    // ---- matchers: tables laid out as assemble_image expects them
    const size_t LC = (size_t)L * capP;
    double *s_cs = reinterpret_cast<double *>(smem_raw);
    double *s_cn = s_cs + LC;
    uint32_t *s_cij = reinterpret_cast<uint32_t *>(s_cn + LC);
    int *s_cc = reinterpret_cast<int *>(s_cij + LC);
    const int n_match = (int)blockDim.x / 32 - 1;
    for (int k = warp - 1; k < L; k += n_match) {
        uint32_t *o_ij = s_cij + (size_t)k * capP;
        double *o_sc = s_cs + (size_t)k * capP, *o_nm = s_cn + (size_t)k * capP;
        SPG_TR(16 + 4 * k, 0);
        const int m = match_limb(ws, n, k, lane, keys_valid != 0, o_ij, o_sc, o_nm);
        __syncwarp();
        SPG_TR(16 + 4 * k + 2, m);
        // the stage-wise API (spg_download_connections, spg_assemble) reads the tables from global memory
        const size_t obase = ((size_t)n * L + k) * capP;
        for (int c = lane; c < m; c += 32) {
            ws.conn_ij[obase + c] = o_ij[c];
            ws.conn_score[obase + c] = o_sc[c];
            ws.conn_norm[obase + c] = o_nm[c];
        }
        __syncwarp();
        if (lane == 0) {
            ws.conn_count[(size_t)n * L + k] = m;
            s_cc[k] = m;
            asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(smem_u32(s_ready + k)), "r"(1) : "memory");
        }
        SPG_TR(16 + 4 * k + 3, 0);
    }
}

}  // namespace spg
