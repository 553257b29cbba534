// dense_tc2.cu -- tensor-core scan, second generation: query block RESIDENT IN TENSOR MEMORY + corpus tiles MULTICAST
// across a thread-block cluster.
//
// Why (profiles/r1a_scan_v1.md): at batch 1024 the first kernel re-streamed the 128x768 query tile for every corpus
// tile and every query block fetched its own copy of each corpus tile; the scan ran at the L2->SM bandwidth limit
// (~7 TB/s) and DRAM saw the corpus 3.4 times.  Here
//   * the 128-query block is written ONCE per work item into TMEM (dim/2 columns, two 16-bit values per column) and
//     is the A operand of tcgen05.mma (TS form), so shared memory carries corpus rows only;
//   * a cluster of CS CTAs holds CS different query blocks and shares every corpus tile: each CTA fetches 1/CS of the
//     stage with cp.async.bulk.tensor ... .multicast::cluster, the stage lands in all CS shared memories, and the
//     tcgen05.commit that frees a stage is multicast to all CS producers.  L2->SM traffic drops from
//     (1 + 1/2) * nqb * |X| to nqb/CS * |X|.
//   * accumulators are N_ACC = 64 (dim <= 768) or 128 (dim <= 512) columns, double buffered in the rest of TMEM.
//
//   warp 0      TMA producer (one elected lane)        warp 1   MMA issuer + TMEM allocator
//   warps 2..5  epilogue: load Q block -> TMEM (tcgen05.st), then per tile tcgen05.ld + threshold filter + candidate
//               append, warp-cooperative register-resident compaction, shared per-query threshold (atomicMax)
#include "tc_common.cuh"

namespace b200rag {

constexpr int T2_THREADS = 192;
constexpr int T2_EPI_WARPS = 4;
constexpr int T2_KB_ROW_BYTES = 128;          // one k-block row: 64 x 16-bit
constexpr int T2_SMEM_STAGE_BUDGET = 196608;  // bytes of shared memory given to the stage ring

template <int N_ACC> struct T2Cfg {
    static constexpr int KB_BYTES = N_ACC * T2_KB_ROW_BYTES;          // one k-block of a tile: 8 KB / 16 KB
    static constexpr int KPS = 32768 / KB_BYTES;                      // k-blocks per stage: 4 / 2  (32 KB stages)
    static constexpr int STAGE_BYTES = KPS * KB_BYTES;
    static constexpr int STAGES = T2_SMEM_STAGE_BUDGET / STAGE_BYTES; // 6
    static constexpr int ACC_COL0 = 512 - 2 * N_ACC;                  // accumulators sit at the top of TMEM
    static constexpr int MAX_KBLOCKS = ACC_COL0 / 32;                 // Q needs 32 columns per k-block
};

template <int N_ACC>
__global__ void __launch_bounds__(T2_THREADS, 1)
dense_scan2_kernel(const __grid_constant__ CUtensorMap map_x, const ScanParams p, const int cs) {
    using Cfg = T2Cfg<N_ACC>;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stage_base = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
    uint64_t* full_bar = bars;                               // [STAGES]   1 arrival (local producer) + tx bytes
    uint64_t* empty_bar = bars + Cfg::STAGES;                // [STAGES]   cs arrivals (every CTA's MMA warp)
    uint64_t* tfull_bar = bars + 2 * Cfg::STAGES;            // [2]
    uint64_t* tempty_bar = bars + 2 * Cfg::STAGES + 2;       // [2]        4 arrivals (epilogue warps)
    uint64_t* qready_bar = bars + 2 * Cfg::STAGES + 4;       // [1]        4 arrivals (epilogue warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 5);
    uint32_t* scratch_all = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::STAGES + 6);   // [4][cap] (large-k compaction)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int rank = cs > 1 ? (int)cluster_ctarank() : 0;
    const int cluster_id = blockIdx.x / cs;
    const int n_clusters = gridDim.x / cs;
    const int qgroups = p.nqb / cs;
    const uint16_t mc_mask = (uint16_t)((1u << cs) - 1);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], cs); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], T2_EPI_WARPS); }
        mbar_init(qready_bar, T2_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    if (cs > 1) cluster_sync_all();          // remote CTAs must not touch our barriers before they are initialised
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int stages_per_tile = (p.n_kblocks + Cfg::KPS - 1) / Cfg::KPS;

    if (warp == 0) {
        // ================================================================= TMA producer
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            long long st_wait = 0;
            ST_T0(st_begin);
            for (int item = cluster_id; item < p.n_items; item += n_clusters) {
                const int chunk = item / qgroups;
                const int t0 = (int)((int64_t)chunk * p.n_tiles / p.n_chunks);
                const int t1 = (int)((int64_t)(chunk + 1) * p.n_tiles / p.n_chunks);
                for (int tile = t0; tile < t1; ++tile) {
                    for (int sg = 0; sg < stages_per_tile; ++sg) {
                        const int kb0 = sg * Cfg::KPS;
                        const int nkb = min(Cfg::KPS, p.n_kblocks - kb0);
                        ST_T0(tw);
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        ST_ADD(st_wait, tw);      // slot free in EVERY CTA of the cluster
                        uint8_t* sx = stage_base + stage * Cfg::STAGE_BYTES;
                        mbar_expect_tx(&full_bar[stage], (uint32_t)(nkb * Cfg::KB_BYTES));
                        for (int j = 0; j < nkb; ++j) {
                            const int kb = kb0 + j;
                            if (cs == 1) {
                                tma_load_2d(sx + j * Cfg::KB_BYTES, &map_x, kb * TC_BK, tile * p.tile_stride * N_ACC, &full_bar[stage]);
                            } else if (kb % cs == rank) {
                                tma_load_2d_mc(sx + j * Cfg::KB_BYTES, &map_x, kb * TC_BK, tile * p.tile_stride * N_ACC, &full_bar[stage], mc_mask);
                            }
                        }
                        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
            if (p.stats) {
                p.stats[blockIdx.x * ST_N + ST_PROD_TOTAL] = clock64() - st_begin;
                p.stats[blockIdx.x * ST_N + ST_PROD_WAIT_EMPTY] = st_wait;
            }
        }
    } else if (warp == 1) {
        // ================================================================= MMA issuer
        int stage = 0, astage = 0;
        uint32_t phase = 0, aphase = 0, qphase = 0;
        const uint32_t idesc = p.idesc;
        long long st_wfull = 0, st_wtempty = 0, st_wq = 0;
        ST_T0(st_begin);
        for (int item = cluster_id; item < p.n_items; item += n_clusters) {
            const int chunk = item / qgroups;
            const int t0 = (int)((int64_t)chunk * p.n_tiles / p.n_chunks);
            const int t1 = (int)((int64_t)(chunk + 1) * p.n_tiles / p.n_chunks);
            ST_T0(tq);
            mbar_wait(qready_bar, qphase);                 // this item's query block is in TMEM
            ST_ADD(st_wq, tq);
            qphase ^= 1;
            tc_fence_after();
            for (int tile = t0; tile < t1; ++tile) {
                ST_T0(te);
                mbar_wait(&tempty_bar[astage], aphase ^ 1);
                ST_ADD(st_wtempty, te);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(Cfg::ACC_COL0 + astage * N_ACC);
                for (int sg = 0; sg < stages_per_tile; ++sg) {
                    const int kb0 = sg * Cfg::KPS;
                    const int nkb = min(Cfg::KPS, p.n_kblocks - kb0);
                    ST_T0(tf);
                    mbar_wait(&full_bar[stage], phase);
                    ST_ADD(st_wfull, tf);
                    tc_fence_after();
                    if (elect_one()) {
                        // descriptors differ only in the 14-bit start-address field: one add per MMA
                        const uint32_t sx = smem_u32(stage_base + stage * Cfg::STAGE_BYTES);
                        const uint32_t desc_lo0 = ((sx & 0x3FFFF) >> 4) | (1u << 16);
                        const uint64_t desc_hi = (uint64_t)((1024u >> 4) | (1u << 14) | (2u << 29)) << 32;
                        const uint32_t a0 = tmem_base + (uint32_t)(kb0 * 32);
#pragma unroll
                        for (int j = 0; j < Cfg::KPS; ++j) {
                            if (j < nkb) {
#pragma unroll
                                for (int k4 = 0; k4 < TC_BK / 16; ++k4) {
                                    const uint32_t acc = (j | k4) != 0 ? 1u : (uint32_t)(sg != 0);
                                    umma_f16_ts(d_tmem, a0 + (uint32_t)((j * 4 + k4) * 8),
                                                desc_hi | (uint64_t)(desc_lo0 + (uint32_t)((j * Cfg::KB_BYTES + k4 * 32) >> 4)), idesc, acc);
                                }
                            }
                        }
                        if (cs == 1) umma_commit(&empty_bar[stage]);
                        else umma_commit_mc(&empty_bar[stage], mc_mask);       // frees the slot for all CS producers
                        if (sg == stages_per_tile - 1) umma_commit(&tfull_bar[astage]);
                    }
                    __syncwarp();
                    if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
                }
                if (++astage == 2) { astage = 0; aphase ^= 1; }
            }
        }
        if (p.stats && lane == 0) {
            p.stats[blockIdx.x * ST_N + ST_MMA_TOTAL] = clock64() - st_begin;
            p.stats[blockIdx.x * ST_N + ST_MMA_WAIT_FULL] = st_wfull;
            p.stats[blockIdx.x * ST_N + ST_MMA_WAIT_TEMPTY] = st_wtempty;
            p.stats[blockIdx.x * ST_N + ST_MMA_WAIT_Q] = st_wq;
        }
    } else {
        // ================================================================= epilogue
        const int quarter = warp & 3;
        const int qlane = quarter * 32 + lane;
        const uint32_t lane_taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        const uint32_t scratch = smem_u32(scratch_all + (size_t)(warp - 2) * p.cap);
        int astage = 0;
        uint32_t aphase = 0;
        long long st_wtfull = 0, st_qload = 0;
        EpiCounters ec;
        ST_T0(st_begin);
        for (int item = cluster_id; item < p.n_items; item += n_clusters) {
            const int chunk = item / qgroups, qg = item % qgroups;
            const int qb = qg * cs + rank;
            const int t0 = (int)((int64_t)chunk * p.n_tiles / p.n_chunks);
            const int t1 = (int)((int64_t)(chunk + 1) * p.n_tiles / p.n_chunks);
            const int q = qb * TC_BM + qlane;
            const bool active = q < p.n_q;
            // ---- query block -> TMEM.  The previous item's MMAs have all retired (its last tfull was observed).
            ST_T0(tql);
            {
                const uint4* qrow = reinterpret_cast<const uint4*>(p.queries + (size_t)(active ? q : 0) * p.dim);
                const int n_vec = p.dim / 8;                              // 16-byte vectors in the row
                for (int kb = 0; kb < p.n_kblocks; ++kb) {
                    uint32_t r[32];
#pragma unroll
                    for (int v = 0; v < 8; ++v) {
                        const int vi = kb * 8 + v;
                        uint4 val = make_uint4(0u, 0u, 0u, 0u);
                        if (active && vi < n_vec) val = __ldg(qrow + vi);
                        r[4 * v + 0] = val.x; r[4 * v + 1] = val.y; r[4 * v + 2] = val.z; r[4 * v + 3] = val.w;
                    }
                    tmem_st32(lane_taddr + (uint32_t)(kb * 32), r);
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(qready_bar);
            }
            ST_ADD(st_qload, tql);
            unsigned long long* buf = p.cand + ((size_t)(chunk * p.nqb + qb) * TC_BM + qlane) * p.cap;
            unsigned int* my_gthr = p.gthr + q;
            float thr = active ? gthr_load(my_gthr) : CUDART_INF_F;
            int cnt = 0;
            float top[TC_SAMPLE_R];
#pragma unroll
            for (int i = 0; i < TC_SAMPLE_R; ++i) top[i] = -CUDART_INF_F;
            for (int tile = t0; tile < t1; ++tile) {
                if (!p.sample && active && ((tile - t0) & 15) == 15) thr = fmaxf(thr, gthr_load(my_gthr));
                ST_T0(tt);
                mbar_wait(&tfull_bar[astage], aphase);
                ST_ADD(st_wtfull, tt);
                tc_fence_after();
                const int64_t row0 = (int64_t)tile * p.tile_stride * N_ACC;
                const uint32_t taddr = lane_taddr + (uint32_t)(Cfg::ACC_COL0 + astage * N_ACC);
                if (p.sample) epi_sample_tile(taddr, N_ACC / 32, row0, p.n_rows, top);
                else epi_filter_tile<N_ACC / 32>(taddr, row0, p.n_rows, thr, cnt, buf, my_gthr, p.kprime, p.cap, scratch, lane, ec);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[astage]);
                if (++astage == 2) { astage = 0; aphase ^= 1; }
            }
            if (p.sample) {
                epi_sample_finish(top, buf);
                cnt = TC_SAMPLE_R;
            } else {
                epi_filter_finish(cnt, buf, my_gthr, p.kprime, p.cap, scratch, lane);
            }
            p.cand_cnt[(size_t)(chunk * p.nqb + qb) * TC_BM + qlane] = cnt;
        }
        if (p.stats && warp == 2 && lane == 0) {
            p.stats[blockIdx.x * ST_N + ST_EPI_TOTAL] = clock64() - st_begin;
            p.stats[blockIdx.x * ST_N + ST_EPI_WAIT_TFULL] = st_wtfull;
            p.stats[blockIdx.x * ST_N + ST_EPI_COMPACT] = ec.compact;
            p.stats[blockIdx.x * ST_N + ST_EPI_QLOAD] = st_qload;
            p.stats[blockIdx.x * ST_N + ST_EPI_NCOMPACT] = ec.ncompact;
            p.stats[blockIdx.x * ST_N + ST_EPI_NSLOW] = ec.nslow;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (cs > 1) cluster_sync_all();          // nobody exits while a peer may still multicast into it
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ----------------------------------------------------------------------------------------------- host side
int make_tensor_map(CUtensorMap* map, const void* base, int64_t rows, int dim, int dtype, int box_rows);

bool scan2_supported(int dim) { return (dim + TC_BK - 1) / TC_BK <= T2Cfg<64>::MAX_KBLOCKS; }
int scan2_tile_rows(int dim) { return (dim + TC_BK - 1) / TC_BK <= T2Cfg<128>::MAX_KBLOCKS ? 128 : 64; }
size_t scan2_smem_bytes(int cap) { return 1024 + (size_t)T2_SMEM_STAGE_BUDGET + (2 * 6 + 6) * 8 + (size_t)T2_EPI_WARPS * cap * 4 + 64; }

template <int N_ACC>
static int launch_scan2_t(const void* corpus16, int dtype, const ScanParams& sp, int cs, int max_ctas, cudaStream_t st, int* grid_out) {
    CUtensorMap map_x;
    int rc = make_tensor_map(&map_x, corpus16, sp.n_rows, sp.dim, dtype, N_ACC);
    if (rc) return rc;
    const size_t smem = scan2_smem_bytes(sp.cap);
    auto kern = dense_scan2_kernel<N_ACC>;
    B200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(T2_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n_clusters = max_ctas / cs;
    if (cs > 1) {
        cfg.gridDim = dim3(max_ctas / cs * cs);
        int max_active = 0;
        if (cudaOccupancyMaxActiveClusters(&max_active, kern, &cfg) == cudaSuccess && max_active > 0 && max_active < n_clusters)
            n_clusters = max_active;          // persistent kernel: every cluster must be co-resident
    }
    if (n_clusters > sp.n_items) n_clusters = sp.n_items;
    if (n_clusters < 1) n_clusters = 1;
    cfg.gridDim = dim3(n_clusters * cs);
    if (grid_out) *grid_out = n_clusters * cs;
    B200_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, map_x, sp, cs));
    count_launch();
    return B200RAG_OK;
}

int launch_scan2(const void* corpus16, int dtype, const ScanParams& sp, int cs, int max_ctas, cudaStream_t st, int* grid_out) {
    if (scan2_tile_rows(sp.dim) == 128) return launch_scan2_t<128>(corpus16, dtype, sp, cs, max_ctas, st, grid_out);
    return launch_scan2_t<64>(corpus16, dtype, sp, cs, max_ctas, st, grid_out);
}

// Number of clusters of `cs` CTAs that can be co-resident (used by the planner to balance chunks).
int scan2_max_clusters(int dim, int cap, int cs, int sm_count) {
    if (cs == 1) return sm_count;
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(T2_THREADS);
    cfg.dynamicSmemBytes = scan2_smem_bytes(cap);
    cfg.gridDim = dim3(sm_count / cs * cs);
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int max_active = 0;
    cudaError_t e;
    if (scan2_tile_rows(dim) == 128) {
        cudaFuncSetAttribute(dense_scan2_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes);
        e = cudaOccupancyMaxActiveClusters(&max_active, dense_scan2_kernel<128>, &cfg);
    } else {
        cudaFuncSetAttribute(dense_scan2_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes);
        e = cudaOccupancyMaxActiveClusters(&max_active, dense_scan2_kernel<64>, &cfg);
    }
    if (e != cudaSuccess || max_active <= 0) { cudaGetLastError(); return sm_count / cs; }
    return max_active < sm_count / cs ? max_active : sm_count / cs;
}

}  // namespace b200rag
