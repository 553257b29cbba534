// dense_tc3.cu -- tensor-core scan, third generation: CTA PAIRS (tcgen05.mma.cta_group::2).
//
// Why (profiles/r1_scan3_ncu.md, history table): the single-CTA kernel moves 48 KB from L2 into shared memory for every
// 128 x 256 x 64 block of multiply-adds (a 16 KB query tile + a 32 KB corpus tile) -- 12x the corpus size per scan at
// batch 1024, ~11 TB/s of L2->SM traffic -- and the board sits on its power cap at ~0.95 GHz.  Two SMs of one TPC
// can issue ONE MMA of M = 256 (two query blocks) x N = 256 (corpus rows) in which each CTA supplies its own 128
// query rows and only HALF of the corpus tile; the tensor cores read the other half from the peer's shared memory.
// Per CTA and k-block that is 16 KB + 16 KB instead of 16 KB + 32 KB for the same amount of math: one third less
// L2->SM traffic and shared-memory fill, one third less operand fetch per MMA.
//
//   cluster = 2 CTAs (rank 0 = leader).  Work item = (chunk of consecutive corpus tiles, SPAN of query-block pairs).
//   Tile-major order: for every corpus tile of its chunk the cluster runs ALL pairs of the span back to back, so the
//   tile comes from HBM once and is re-served from L2 microseconds later for the other pairs (r1d profile: with one pair
//   per item the 4 clusters sharing a chunk drifted apart and HBM saw the corpus 2.0 times).
//   warp 0      TMA producer in BOTH CTAs: Q tile [128 x 64] of the CTA's own query block + X half tile [128 x 64]
//               (rows tile*256 + rank*128 ...), 6-stage ring of 32 KB; every load credits its bytes to the LEADER's
//               full barrier (cp.async.bulk.tensor ... .cta_group::2)
//   warp 1      MMA issuer, leader only: tcgen05.mma.cta_group::2.kind::f16 M=256 N=256 K=16; tcgen05.commit
//               multicast to both CTAs frees the smem stage in each and publishes the accumulator to each epilogue
//   warps 2..5  epilogue in BOTH CTAs (identical to the first generation): tcgen05.ld 32 columns at a time, a thread
//               owns one query, threshold filter, append of the rare survivors, warp-cooperative compaction.  The
//               accumulator is handed back by arriving on the LEADER's tmem-empty barrier (8 arrivals: 4 warps x 2 CTAs).
#include <cstdlib>
#include <mutex>

#include "tc_common.cuh"

namespace b200rag {

constexpr int T3_THREADS = 192;
constexpr int T3_EPI_WARPS = 4;
constexpr int T3_BN = 256;                               // corpus rows per accumulator (UMMA N)
constexpr int T3_HALF = T3_BN / 2;                       // rows of it staged by each CTA
constexpr int T3_Q_BYTES = TC_BM * TC_BK * 2;            // 16 KB
constexpr int T3_X_BYTES = T3_HALF * TC_BK * 2;          // 16 KB
constexpr int T3_STAGE_BYTES = T3_Q_BYTES + T3_X_BYTES;  // 32 KB
constexpr int T3_STAGES = 6;
constexpr int T3_N_BARS = 2 * T3_STAGES + 4;

template <int EPI_VAR>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T3_THREADS, 1)
dense_scan3_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x, const ScanParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stage_base = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T3_STAGES * T3_STAGE_BYTES);
    uint64_t* full_bar = bars;                          // [STAGES]  leader: 1 arrival (its producer) + 2 x 32 KB of tx
    uint64_t* empty_bar = bars + T3_STAGES;             // [STAGES]  each CTA: 1 arrival (multicast tcgen05.commit)
    uint64_t* tfull_bar = bars + 2 * T3_STAGES;         // [2]       each CTA: 1 arrival (multicast tcgen05.commit)
    uint64_t* tempty_bar = bars + 2 * T3_STAGES + 2;    // [2]       leader: 8 arrivals (epilogue warps of both CTAs)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + T3_N_BARS);
    uint32_t* scratch_all = reinterpret_cast<uint32_t*>(bars + T3_N_BARS + 2);   // [4][cap] (large-k compaction)
    float* st_thr = reinterpret_cast<float*>(scratch_all + (size_t)T3_EPI_WARPS * p.cap);   // [qg_span][128] running thresholds
    int* st_cnt = reinterpret_cast<int*>(st_thr + (size_t)p.qg_span * TC_BM);                // [qg_span][128] buffer fill

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int rank = (int)cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1;
    const int n_clusters = gridDim.x >> 1;
    const int qgroups_all = p.nqb >> 1;                 // nqb is padded to an even number of query blocks
    const int n_spans = (qgroups_all + p.qg_span - 1) / p.qg_span;
    // device-side gate (tier-0 re-scan of flagged queries): only the first *gate queries exist; pairs of query blocks beyond
    // them are skipped by every role alike, so a launch with nothing to do costs a few microseconds
    const int qgroups = p.gate ? min(qgroups_all, (__ldg(p.gate) + 2 * TC_BM - 1) / (2 * TC_BM)) : qgroups_all;
    const int n_q_live = p.gate ? min(p.n_q, __ldg(p.gate)) : p.n_q;      // (padding rows of the last live pair never pass)

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_q)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
        for (int s = 0; s < T3_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 2 * T3_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {    // one warp of EACH CTA of the pair takes part in the paired allocation
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();             // the peer must not signal our barriers before they are initialised
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================================================= TMA producer (one elected lane, both CTAs)
        if (elect_one()) {
            const uint32_t full_leader = mapa_u32(smem_u32(full_bar), 0);
            int stage = 0;
            uint32_t phase = 0;
            long long st_wait = 0;
            ST_T0(st_begin);
            for (int item = cluster_id; item < p.n_items; item += n_clusters) {
                const int chunk = item / n_spans, qg0 = (item % n_spans) * p.qg_span;
                const int qg1 = max(qg0, min(qg0 + p.qg_span, qgroups));
                const int t0 = (int)((int64_t)chunk * p.n_tiles / p.n_chunks);
                const int t1 = (int)((int64_t)(chunk + 1) * p.n_tiles / p.n_chunks);
                if (qg1 == qg0) continue;
                for (int tile = t0; tile < t1; ++tile) {
                    const int xrow = tile * p.tile_stride * T3_BN + rank * T3_HALF;
                    for (int qg = qg0; qg < qg1; ++qg) {
                        const int qrow = (qg * 2 + rank) * TC_BM;
                        for (int kb = 0; kb < p.n_kblocks; ++kb) {
                            ST_T0(tw);
                            mbar_wait(&empty_bar[stage], phase ^ 1);      // the pair's MMAs have drained OUR copy of the slot
                            ST_ADD(st_wait, tw);
                            uint8_t* sq = stage_base + stage * T3_STAGE_BYTES;
                            uint8_t* sx = sq + T3_Q_BYTES;
                            if (leader) mbar_expect_tx(&full_bar[stage], 2 * T3_STAGE_BYTES);
                            tma_load_2d_pair(sq, &map_q, kb * TC_BK, qrow, full_leader + stage * 8);
                            tma_load_2d_pair(sx, &map_x, kb * TC_BK, xrow, full_leader + stage * 8);
                            if (++stage == T3_STAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
            if (p.stats) {
                p.stats[blockIdx.x * ST_N + ST_PROD_TOTAL] = clock64() - st_begin;
                p.stats[blockIdx.x * ST_N + ST_PROD_WAIT_EMPTY] = st_wait;
            }
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer (leader CTA only)
        if (leader) {
            int stage = 0, astage = 0;
            uint32_t phase = 0, aphase = 0;
            long long st_wfull = 0, st_wtempty = 0;
            ST_T0(st_begin);
            for (int item = cluster_id; item < p.n_items; item += n_clusters) {
                const int chunk = item / n_spans, qg0 = (item % n_spans) * p.qg_span;
                const int n_acc = max(0, min(qg0 + p.qg_span, qgroups) - qg0) *
                                  ((int)((int64_t)(chunk + 1) * p.n_tiles / p.n_chunks) - (int)((int64_t)chunk * p.n_tiles / p.n_chunks));
                for (int acc = 0; acc < n_acc; ++acc) {       // one accumulator per (tile, pair of query blocks)
                    ST_T0(te);
                    mbar_wait_cluster(&tempty_bar[astage], aphase ^ 1);   // both epilogues have drained this accumulator
                    ST_ADD(st_wtempty, te);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + (uint32_t)astage * T3_BN;
                    for (int kb = 0; kb < p.n_kblocks; ++kb) {
                        ST_T0(tf);
                        mbar_wait_cluster(&full_bar[stage], phase);       // both CTAs' TMA bytes have landed
                        ST_ADD(st_wfull, tf);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t sq = smem_u32(stage_base + stage * T3_STAGE_BYTES);
                            const uint32_t sx = sq + T3_Q_BYTES;
#pragma unroll
                            for (int k4 = 0; k4 < TC_BK / 16; ++k4) {
                                umma_f16_ss_pair(d_tmem, umma_desc_sw128(sq + k4 * 32), umma_desc_sw128(sx + k4 * 32), p.idesc,
                                                 (uint32_t)((kb | k4) != 0));
                            }
                            umma_commit_pair(&empty_bar[stage], 3);                               // slot reusable in both CTAs
                            if (kb == p.n_kblocks - 1) umma_commit_pair(&tfull_bar[astage], 3);   // accumulator complete
                        }
                        __syncwarp();
                        if (++stage == T3_STAGES) { stage = 0; phase ^= 1; }
                    }
                    if (++astage == 2) { astage = 0; aphase ^= 1; }
                }
            }
            if (p.stats && lane == 0) {
                p.stats[blockIdx.x * ST_N + ST_MMA_TOTAL] = clock64() - st_begin;
                p.stats[blockIdx.x * ST_N + ST_MMA_WAIT_FULL] = st_wfull;
                p.stats[blockIdx.x * ST_N + ST_MMA_WAIT_TEMPTY] = st_wtempty;
            }
        }
    } else {
        // ================================================================= epilogue: fused threshold filter (both CTAs)
        const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
        const int qlane = quarter * 32 + lane;        // query row inside the block == TMEM lane
        const uint32_t scratch = smem_u32(scratch_all + (size_t)(warp - 2) * p.cap);
        const uint32_t tempty_leader = mapa_u32(smem_u32(tempty_bar), 0);
        int astage = 0;
        uint32_t aphase = 0;
        long long st_wtfull = 0;
        EpiCounters ec;
        ST_T0(st_begin);
        for (int item = cluster_id; item < p.n_items; item += n_clusters) {
            const int chunk = item / n_spans, qg0 = (item % n_spans) * p.qg_span;
            const int qg1 = max(qg0, min(qg0 + p.qg_span, qgroups));
            const int t0 = (int)((int64_t)chunk * p.n_tiles / p.n_chunks);
            const int t1 = (int)((int64_t)(chunk + 1) * p.n_tiles / p.n_chunks);
            if (qg1 == qg0) continue;
            if (p.sample) {
                // ---- SAMPLE pass (qg_span == 1): best group maxima in registers
                const int qb = qg0 * 2 + rank;
                const size_t slot = (size_t)(chunk * p.nqb + qb) * TC_BM + qlane;
                float top[TC_SAMPLE_R];
#pragma unroll
                for (int i = 0; i < TC_SAMPLE_R; ++i) top[i] = -CUDART_INF_F;
                for (int tile = t0; tile < t1; ++tile) {
                    ST_T0(tt);
                    mbar_wait(&tfull_bar[astage], aphase);
                    ST_ADD(st_wtfull, tt);
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)astage * T3_BN;
                    epi_sample_tile(taddr, T3_BN / 32, (int64_t)tile * p.tile_stride * T3_BN, p.n_rows, top, p.row_mask);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(tempty_leader + astage * 8);
                    if (++astage == 2) { astage = 0; aphase ^= 1; }
                }
                epi_sample_finish(top, p.cand + slot * p.cap);
                p.cand_cnt[slot] = TC_SAMPLE_R;
                continue;
            }
            // ---- FULL scan: per-query state of every pair of the span lives in shared memory (a thread only ever touches
            //      its own slots, so no synchronisation is needed)
            for (int qg = qg0; qg < qg1; ++qg) {
                const int q = (qg * 2 + rank) * TC_BM + qlane;
                // start from the best threshold any CTA has established for this query so far (valid lower bound of the
                // global k'-th best score; -inf while nobody has k' candidates yet); padding lanes never pass
                st_thr[(qg - qg0) * TC_BM + qlane] = q < n_q_live ? gthr_load(p.gthr + q) : CUDART_INF_F;
                st_cnt[(qg - qg0) * TC_BM + qlane] = 0;
            }
            for (int tile = t0; tile < t1; ++tile) {
                const int64_t row0 = (int64_t)tile * p.tile_stride * T3_BN;
                for (int qg = qg0; qg < qg1; ++qg) {
                    const int qb = qg * 2 + rank;
                    const int q = qb * TC_BM + qlane;
                    const int si = (qg - qg0) * TC_BM + qlane;
                    unsigned long long* buf = p.cand + ((size_t)(chunk * p.nqb + qb) * TC_BM + qlane) * p.cap;
                    unsigned int* my_gthr = p.gthr + q;       // gthr has nqb*128 entries, padded blocks included
                    float thr = st_thr[si];
                    int cnt = st_cnt[si];
                    if (((tile - t0) & 7) == 7 && q < n_q_live) thr = fmaxf(thr, gthr_load(my_gthr));
                    ST_T0(tt);
                    mbar_wait(&tfull_bar[astage], aphase);
                    ST_ADD(st_wtfull, tt);
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)astage * T3_BN;
                    epi_filter_tile<T3_BN / 32, EPI_VAR>(taddr, row0, p.n_rows, thr, cnt, buf, my_gthr, p.kprime, p.cap, scratch, lane, ec, p.row_mask);
                    // accumulator drained: hand it back to the leader's MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(tempty_leader + astage * 8);
                    if (++astage == 2) { astage = 0; aphase ^= 1; }
                    st_thr[si] = thr;
                    st_cnt[si] = cnt;
                }
            }
            for (int qg = qg0; qg < qg1; ++qg) {
                const int qb = qg * 2 + rank;
                const size_t slot = (size_t)(chunk * p.nqb + qb) * TC_BM + qlane;
                int cnt = st_cnt[(qg - qg0) * TC_BM + qlane];
                epi_filter_finish(cnt, p.cand + slot * p.cap, p.gthr + qb * TC_BM + qlane, p.kprime, p.cap, scratch, lane);
                p.cand_cnt[slot] = cnt;
            }
        }
        if (p.stats && warp == 2 && lane == 0) {
            p.stats[blockIdx.x * ST_N + ST_EPI_TOTAL] = clock64() - st_begin;
            p.stats[blockIdx.x * ST_N + ST_EPI_WAIT_TFULL] = st_wtfull;
            p.stats[blockIdx.x * ST_N + ST_EPI_COMPACT] = ec.compact;
            p.stats[blockIdx.x * ST_N + ST_EPI_NCOMPACT] = ec.ncompact;
            p.stats[blockIdx.x * ST_N + ST_EPI_NSLOW] = ec.nslow;
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();             // nobody frees tensor memory or exits while the peer may still use it
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS));
    }
}

// ----------------------------------------------------------------------------------------------- host side
int make_tensor_map(CUtensorMap* map, const void* base, int64_t rows, int dim, int dtype, int box_rows);

size_t scan3_smem_bytes(int cap, int span) {
    return 1024 + (size_t)T3_STAGES * T3_STAGE_BYTES + (T3_N_BARS + 2) * 8 + (size_t)T3_EPI_WARPS * cap * 4 +
           (size_t)span * TC_BM * 8 + 64;
}

int scan3_max_clusters_query(int cap, int span, int sm_count);
// Number of CTA pairs that can be co-resident (the kernel is persistent: every pair must be resident at once).
int scan3_max_clusters(int cap, int span, int sm_count) {
    // the occupancy query costs tens of microseconds per call: one cached answer, guarded (the C ABI is reentrant)
    static std::mutex mu;
    struct Entry { int cap, span, sm, val; };
    static Entry cache[8] = {};                         // (the first pass and the tier-0 re-scan alternate between two shapes)
    static int n_cached = 0, next = 0;
    std::lock_guard<std::mutex> lock(mu);
    for (int i = 0; i < n_cached; ++i)
        if (cache[i].cap == cap && cache[i].span == span && cache[i].sm == sm_count) return cache[i].val;
    const int val = scan3_max_clusters_query(cap, span, sm_count);
    cache[next] = Entry{cap, span, sm_count, val};
    next = (next + 1) % 8;
    if (n_cached < 8) ++n_cached;
    return val;
}

int scan3_max_clusters_query(int cap, int span, int sm_count) {
    const size_t smem = scan3_smem_bytes(cap, span);
    if (cudaFuncSetAttribute(dense_scan3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return sm_count / 2;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(T3_THREADS);
    cfg.gridDim = dim3(sm_count / 2 * 2);
    cfg.dynamicSmemBytes = smem;
    int max_active = 0;
    if (cudaOccupancyMaxActiveClusters(&max_active, dense_scan3_kernel<1>, &cfg) != cudaSuccess || max_active <= 0) {
        cudaGetLastError();
        return sm_count / 2;
    }
    return max_active < sm_count / 2 ? max_active : sm_count / 2;
}

int launch_scan3(const void* corpus16, int dtype, const ScanParams& sp, int max_clusters, int span_cap, int span_max, cudaStream_t st) {
    // span_cap / span_max: the FULL scan's buffer capacity and span, so that the sample pass is launched with the same
    // shared-memory size (and therefore the same co-residency) as the full scan
    CUtensorMap map_q, map_x;
    int rc = make_tensor_map(&map_q, sp.queries, (int64_t)sp.nqb * TC_BM, sp.dim, dtype, TC_BM);   // whole blocks (padded by run_tensor)
    if (rc) return rc;
    rc = make_tensor_map(&map_x, corpus16, sp.n_rows, sp.dim, dtype, T3_HALF);
    if (rc) return rc;
    const size_t smem = scan3_smem_bytes(sp.cap > span_cap ? sp.cap : span_cap, span_max);
    const bool var0 = option(OPT_EPI, 1) == 0;             // A/B switch of the epilogue variant (tc_common.cuh: epi_filter_group)
    if (var0) B200_CUDA_CHECK(cudaFuncSetAttribute(dense_scan3_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else B200_CUDA_CHECK(cudaFuncSetAttribute(dense_scan3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int n_clusters = sp.n_items < max_clusters ? sp.n_items : max_clusters;
    if (n_clusters < 1) n_clusters = 1;
    if (var0) dense_scan3_kernel<0><<<n_clusters * 2, T3_THREADS, smem, st>>>(map_q, map_x, sp);
    else dense_scan3_kernel<1><<<n_clusters * 2, T3_THREADS, smem, st>>>(map_q, map_x, sp);
    count_launch();
    B200_CUDA_CHECK(cudaGetLastError());
    return B200RAG_OK;
}

}  // namespace b200rag
