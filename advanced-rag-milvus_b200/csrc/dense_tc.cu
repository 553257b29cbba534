// dense_tc.cu -- the tensor-core flat scan with the top-k filter fused into the epilogue  (K1 + K2).
//
//   scan kernel (persistent, one CTA per SM, warp specialised, sm_100a only)
//     warp 0      TMA producer: cp.async.bulk.tensor tiles of Q [128 x 64] and X [256 x 64] (SWIZZLE_128B) into a
//                 4-stage shared-memory ring, mbarrier complete_tx signalling
//     warp 1      MMA issuer: tcgen05.mma.cta_group::1.kind::f16, M=128 (queries) x N=256 (corpus rows) x K=16,
//                 fp32 accumulators double-buffered in TMEM (2 x 256 columns); tcgen05.commit frees smem stages and
//                 publishes finished accumulators
//     warps 2..5  epilogue: tcgen05.ld 32 columns at a time; a thread owns ONE query (TMEM lane), compares the 32 scores
//                 against that query's running threshold (the k'-th best so far) and appends the rare survivors
//                 (score, row) to a per-query candidate buffer in global memory; a warp-cooperative radix select
//                 compacts a buffer back to k' entries and raises the threshold when it fills.  The score matrix never
//                 leaves the SM.
//     Work item = (query block of 128, chunk of consecutive corpus tiles); concurrently resident CTAs hold different
//     query blocks of the same chunks so the corpus streams from HBM once and is re-served from L2.
//
//   finish kernel (one CTA per query)
//     merges the per-chunk candidate lists to the k' best by tensor-core score, RE-SCORES them in the canonical fp64
//     arithmetic (bit-identical to the oracle), sorts by (score desc, id asc) and PROVES completeness:
//     every row outside the candidate set has tensor-core score <= m (the k'-th best), hence exact score <= m + eps;
//     if the exact k-th score exceeds m + eps the exact top-k is inside the candidate set.  Otherwise the query is
//     flagged and (AUTO mode) re-run on the exact CUDA-core path.  eps = 2 * dim * 2^-23 * |q| * row_norm_bound bounds
//     the fp32 accumulation error of the tensor-core path (checked empirically in tests/test_gpu_dense_tc.py).
#include <cmath>
#include "tc_common.cuh"
#include "select.cuh"
#include "finish.cuh"

namespace b200rag {

size_t exact_workspace_bytes(int64_t n_rows, int dim, int n_q, int k);
int run_exact(const void* corpus16, int64_t n_rows, int dim, int dtype, const void* queries16, int n_launch,
              const int32_t* q_list, int k, int64_t id_offset, double* out_scores, int64_t* out_ids,
              void* workspace, size_t workspace_bytes, cudaStream_t st, const int32_t* n_active, int slot_base,
              const uint32_t* row_mask);

// ----------------------------------------------------------------------------------------------- scan kernel
__global__ void __launch_bounds__(TC_THREADS, 1)
dense_scan_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x, const ScanParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-byte alignment is required by SWIZZLE_128B; the dynamic smem base is only guaranteed 16-byte aligned
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stage_base = smem;                                             // TC_STAGES * TC_STAGE_BYTES
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_STAGES * TC_STAGE_BYTES);
    uint64_t* full_bar = bars;                     // [TC_STAGES]
    uint64_t* empty_bar = bars + TC_STAGES;        // [TC_STAGES]
    uint64_t* tfull_bar = bars + 2 * TC_STAGES;    // [2]
    uint64_t* tempty_bar = bars + 2 * TC_STAGES + 2;   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 4);
    uint32_t* scratch_all = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 6);   // [TC_EPI_WARPS][cap]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_q)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_x)) : "memory");
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], TC_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================================================================= TMA producer (one elected lane)
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            long long st_wait = 0;
            ST_T0(st_begin);
            for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
                const int chunk = item / p.nqb, qb = item % p.nqb;
                const int t0 = (int)((int64_t)chunk * p.n_tiles / p.n_chunks);
                const int t1 = (int)((int64_t)(chunk + 1) * p.n_tiles / p.n_chunks);
                for (int tile = t0; tile < t1; ++tile) {
                    for (int kb = 0; kb < p.n_kblocks; ++kb) {
                        ST_T0(tw);
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        ST_ADD(st_wait, tw);
                        uint8_t* sq = stage_base + stage * TC_STAGE_BYTES;
                        uint8_t* sx = sq + TC_Q_BYTES;
                        mbar_expect_tx(&full_bar[stage], TC_STAGE_BYTES);
                        tma_load_2d(sq, &map_q, kb * TC_BK, qb * TC_BM, &full_bar[stage]);
                        tma_load_2d(sx, &map_x, kb * TC_BK, tile * p.tile_stride * TC_BN, &full_bar[stage]);
                        if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
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
        uint32_t phase = 0, aphase = 0;
        long long st_wfull = 0, st_wtempty = 0, st_wq = 0;
        ST_T0(st_begin);
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            const int chunk = item / p.nqb;
            const int t0 = (int)((int64_t)chunk * p.n_tiles / p.n_chunks);
            const int t1 = (int)((int64_t)(chunk + 1) * p.n_tiles / p.n_chunks);
            for (int tile = t0; tile < t1; ++tile) {
                ST_T0(te);
                mbar_wait(&tempty_bar[astage], aphase ^ 1);      // epilogue has drained this accumulator
                ST_ADD(st_wtempty, te);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)astage * TC_BN;
                for (int kb = 0; kb < p.n_kblocks; ++kb) {
                    ST_T0(tf);
                    mbar_wait(&full_bar[stage], phase);          // TMA bytes have landed
                    ST_ADD(st_wfull, tf);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t sq = smem_u32(stage_base + stage * TC_STAGE_BYTES);
                        const uint32_t sx = sq + TC_Q_BYTES;
#pragma unroll
                        for (int k4 = 0; k4 < TC_BK / 16; ++k4) {
                            umma_f16_ss(d_tmem, umma_desc_sw128(sq + k4 * 32), umma_desc_sw128(sx + k4 * 32), p.idesc,
                                        (uint32_t)((kb | k4) != 0));
                        }
                        umma_commit(&empty_bar[stage]);          // smem stage reusable once these MMAs retire
                        if (kb == p.n_kblocks - 1) umma_commit(&tfull_bar[astage]);   // accumulator complete
                    }
                    __syncwarp();
                    if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
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
        // ================================================================= epilogue: fused threshold filter
        const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
        const int qlane = quarter * 32 + lane;        // query row inside the block == TMEM lane
        const uint32_t scratch = smem_u32(scratch_all + (size_t)(warp - 2) * p.cap);
        int astage = 0;
        uint32_t aphase = 0;
        long long st_wtfull = 0;
        EpiCounters ec;
        ST_T0(st_begin);
        for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
            const int chunk = item / p.nqb, qb = item % p.nqb;
            const int t0 = (int)((int64_t)chunk * p.n_tiles / p.n_chunks);
            const int t1 = (int)((int64_t)(chunk + 1) * p.n_tiles / p.n_chunks);
            const bool active = qb * TC_BM + qlane < p.n_q;
            unsigned long long* buf = p.cand + ((size_t)item * TC_BM + qlane) * p.cap;
            unsigned int* my_gthr = p.gthr + qb * TC_BM + qlane;
            // start from the best threshold any CTA has established for this query so far (valid lower bound of the
            // global k'-th best score; -inf while nobody has k' candidates yet)
            float thr = active ? gthr_load(my_gthr) : CUDART_INF_F;
            int cnt = 0;
            float top[TC_SAMPLE_R];
#pragma unroll
            for (int i = 0; i < TC_SAMPLE_R; ++i) top[i] = -CUDART_INF_F;
            for (int tile = t0; tile < t1; ++tile) {
                if (!p.sample && active && ((tile - t0) & 3) == 3) thr = fmaxf(thr, gthr_load(my_gthr));
                ST_T0(tt);
                mbar_wait(&tfull_bar[astage], aphase);
                ST_ADD(st_wtfull, tt);
                tc_fence_after();
                const int64_t row0 = (int64_t)tile * p.tile_stride * TC_BN;
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)astage * TC_BN;
                if (p.sample) epi_sample_tile(taddr, TC_BN / 32, row0, p.n_rows, top, p.row_mask);
                else epi_filter_tile<TC_BN / 32, 1>(taddr, row0, p.n_rows, thr, cnt, buf, my_gthr, p.kprime, p.cap, scratch, lane, ec, p.row_mask);
                // accumulator drained: hand it back to the MMA warp
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
            p.cand_cnt[(size_t)item * TC_BM + qlane] = cnt;
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
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS));
    }
}

// ----------------------------------------------------------------------------------------------- finish kernel
constexpr int FN_THREADS = 256;

template <int DTYPE>
__global__ void __launch_bounds__(FN_THREADS) dense_finish_kernel(const FinishParams p) {
    extern __shared__ __align__(16) char smem[];
    const int tid = threadIdx.x;
    const int q = blockIdx.x;
    const int qb = q / TC_BM, ql = q % TC_BM;
    double* qd = reinterpret_cast<double*>(smem);                   // [dim]
    double* exact = qd + p.dim;                                     // [kprime]
    uint32_t* rows = reinterpret_cast<uint32_t*>(exact + p.kprime);  // [kprime]
    int* s_cnt = reinterpret_cast<int*>(rows + p.kprime);           // [n_chunks] fill of this query's buffer in every chunk
    char* tkmem = reinterpret_cast<char*>(s_cnt + p.n_chunks);
    tkmem = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(tkmem) + 15) & ~uintptr_t(15));
    __shared__ double s_q2;
    __shared__ float s_err;

    BlockTopK<FN_THREADS, uint32_t> tk;
    char* stage = tk.attach(tkmem, p.topk_cap, p.kprime, FN_THREADS, /*start_digit=*/BlockTopK<FN_THREADS, uint32_t>::NLO + 3);
    tk.init();
    for (int d = tid; d < p.dim; d += FN_THREADS) qd[d] = bits_to_double<DTYPE>(p.queries[(size_t)q * p.dim + d]);
    if (tid == 0) s_err = 0.f;
    __syncthreads();

    // 1. k' best by tensor-core score over all chunks of this query.  Chunk fill counts go to shared memory first; then
    //    warp w takes chunks w, w+8, ... with the NEXT round's entries already in flight while this round's are offered
    //    (one settle per round of 8 chunks x 32 entries; a chunk rarely holds more than a handful of survivors).
    {
        const int warp = tid >> 5, lane = tid & 31;
        constexpr int NW = FN_THREADS / 32;
        for (int c = tid; c < p.n_chunks; c += FN_THREADS) s_cnt[c] = p.cand_cnt[((size_t)(c * p.nqb + qb)) * TC_BM + ql];
        __syncthreads();
        auto entry = [&](int chunk, int i) -> unsigned long long {
            return (chunk < p.n_chunks && i < s_cnt[chunk]) ? __ldg(p.cand + (((size_t)(chunk * p.nqb + qb)) * TC_BM + ql) * p.cap + i) : 0ull;
        };
        unsigned long long e_cur = entry(warp, lane);
        for (int c0 = 0; c0 < p.n_chunks; c0 += NW) {
            const int chunk = c0 + warp;
            const unsigned long long e_next = entry(chunk + NW, lane);
            const int n = chunk < p.n_chunks ? s_cnt[chunk] : 0;
            tk.offer(lane < n, (uint64_t)mono32(__uint_as_float((uint32_t)(e_cur >> 32))), ~(uint32_t)e_cur);
            tk.settle();
            for (int base = 32; __syncthreads_or(base < n); base += 32) {          // chunks with more than 32 survivors
                const unsigned long long e = entry(chunk, base + lane);
                tk.offer(base + lane < n, (uint64_t)mono32(__uint_as_float((uint32_t)(e >> 32))), ~(uint32_t)e);
                tk.settle();
            }
            e_cur = e_next;
        }
    }
    __syncthreads();
    tk.finalize();
    const int n = tk.count();
    const uint64_t* oh = tk.out_hi();
    const uint32_t* ol = tk.out_lo();
    // m: every row outside the candidate set has tensor-core score <= m  (-inf if nothing was ever dropped)
    // (scan kernels drop a row only when its score is <= a threshold that was pushed to gthr; the merge above drops
    //  rows <= the k'-th best of what was emitted)
    const unsigned int gk = p.gthr[q];
    const float m = fmaxf((n >= p.kprime) ? unmono32((uint32_t)oh[n - 1]) : -CUDART_INF_F, gk ? unmono32(gk) : -CUDART_INF_F);

    // 2. exact canonical re-score.  Rows are staged through shared memory 32 at a time with 16-byte loads (every thread
    //    has dim/64 independent loads in flight: one exposed DRAM latency per batch instead of one per element), then
    //    eight threads share a row, one per canonical lane (lane j sums d = j mod 8 in increasing d, exactly as
    //    oracle/exact_scan.c does), and the lanes are combined in the canonical tree with shuffles.
    float my_err = 0.f;
    {
        const int RB = p.stage_rows;                             // rows per batch (<= FN_THREADS / 8)
        const int l8 = tid & 7, grp = tid >> 3;
        const int vec_per_row = p.dim / 8;                        // 16-byte vectors per row
        const int stride = p.dim * 2 + 16;                        // staged row pitch: +16 bytes keeps the 4 rows of a warp on different banks
        const double* qj = qd + l8;
        for (int i0 = 0; i0 < n; i0 += RB) {
            const int rows_here = min(RB, n - i0);
            // cp.async: the copies do not pass through registers, so all of a thread's 16-byte loads are in flight.
            // Warp w copies rows w, w + 8, ...; its lanes stride over the row's 16-byte vectors.
            for (int r = tid >> 5; r < rows_here; r += FN_THREADS / 32) {
                const uint32_t row = ~ol[i0 + r];
                const uint4* src = reinterpret_cast<const uint4*>(p.corpus + (size_t)row * p.dim);
                const uint32_t dst = smem_u32(stage + (size_t)r * stride);
                for (int c = tid & 31; c < vec_per_row; c += 32)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + c * 16), "l"(src + c) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
            const int i = i0 + grp;
            const bool valid = i < n && grp < RB;
            const uint16_t* x = reinterpret_cast<const uint16_t*>(stage + (size_t)grp * stride) + l8;
            double acc = 0.0;
            if (valid) {
#pragma unroll 4
                for (int d = 0; d < p.dim; d += 8) acc = fma(qj[d], bits_to_double<DTYPE>(x[d]), acc);
            }
            double t = __dadd_rn(acc, __shfl_down_sync(0xffffffffu, acc, 1));      // lanes 0,2,4,6: p0+p1, p2+p3, ...
            t = __dadd_rn(t, __shfl_down_sync(0xffffffffu, t, 2));                 // lanes 0,4: (p0+p1)+(p2+p3), ...
            t = __dadd_rn(t, __shfl_down_sync(0xffffffffu, t, 4));                 // lane 0: the canonical score
            if (valid && l8 == 0) {
                rows[i] = ~ol[i];
                exact[i] = t;
                my_err = fmaxf(my_err, fabsf((float)((double)unmono32((uint32_t)oh[i]) - t)));
            }
            __syncthreads();
        }
    }
    if (tid == 0) {
        double s = 0.0;
        for (int d = 0; d < p.dim; ++d) s += qd[d] * qd[d];
        s_q2 = s;
    }
    if (p.err_max) atomicMax(reinterpret_cast<int*>(&s_err), __float_as_int(my_err));   // non-negative floats order as ints
    __syncthreads();
    // 3. rank by (exact desc, row asc) and emit.  Small candidate sets: all-pairs rank count (n^2 / 256 compares per
    //    thread, no barriers); large ones (k' > 256): the selection buffers double as a bitonic sorter.
    const int kk = min(p.k, n);
    __shared__ double s_ek_sh;
    double s_ek = -CUDART_INF;
    if (n <= 256) {
        if (tid == 0) s_ek_sh = -CUDART_INF;
        __syncthreads();
        for (int i = tid; i < n; i += FN_THREADS) {
            const double e = exact[i];
            const uint32_t r = rows[i];
            int rk = 0;
            for (int j = 0; j < n; ++j) rk += (exact[j] > e) || (exact[j] == e && rows[j] < r);
            if (rk < p.k) {
                p.out_scores[(size_t)q * p.k + rk] = e;
                p.out_ids[(size_t)q * p.k + rk] = p.id_offset + (int64_t)r;
            }
            if (rk == kk - 1) s_ek_sh = e;
        }
        for (int i = kk + tid; i < p.k; i += FN_THREADS) {
            p.out_scores[(size_t)q * p.k + i] = -CUDART_INF;
            p.out_ids[(size_t)q * p.k + i] = -1;
        }
        __syncthreads();
        s_ek = s_ek_sh;
    } else {
        tk.init();
        __syncthreads();
        for (int i0 = 0; i0 < n; i0 += FN_THREADS) {              // (offer() wants whole warps)
            const int i = i0 + tid;
            tk.offer(i < n, i < n ? mono64(exact[i]) : 0, i < n ? ~rows[i] : 0);
        }
        __syncthreads();
        tk.finalize();
        const uint64_t* sh = tk.out_hi();
        const uint32_t* sl = tk.out_lo();
        for (int i = tid; i < p.k; i += FN_THREADS) {
            p.out_scores[(size_t)q * p.k + i] = i < kk ? unmono64(sh[i]) : -CUDART_INF;
            p.out_ids[(size_t)q * p.k + i] = i < kk ? p.id_offset + (int64_t)(~sl[i]) : -1;
        }
        if (kk > 0) s_ek = unmono64(sh[kk - 1]);
    }
    __syncthreads();
    if (tid == 0) {
        const double eps = 2.0 * (double)p.dim * 1.1920928955078125e-07 * sqrt(s_q2) * p.row_norm_bound;
        // proven complete iff nothing was dropped (m = -inf) or the k-th exact score clears m + eps
        const bool proven = (m == -CUDART_INF_F) || (n >= p.k && s_ek > (double)m + eps);
        const int flag = proven ? 0 : 1;
        if (p.out_flags) p.out_flags[q] = flag;
        if (flag) p.flag_list[atomicAdd(p.n_flagged, 1)] = q;
        if (p.err_max) p.err_max[q] = s_err;
    }
}

// ----------------------------------------------------------------------------------------------- sample threshold
// After the strided SAMPLE pass: the r-th best tensor-core score among a query's sample candidates becomes its
// starting threshold for the full scan.  With a sample of 1/stride of the rows, about r*stride rows of the whole
// corpus beat it, so r is chosen to leave ~6 k' survivors: enough to contain the top-k' with overwhelming probability,
// few enough that the epilogue's append path and the compaction are rare.  A threshold that turns out too high only
// costs speed (the query fails the completeness proof and is re-run exactly), never correctness.
__global__ void __launch_bounds__(FN_THREADS)
sample_threshold_kernel(const unsigned long long* __restrict__ cand, const int* __restrict__ cand_cnt, int cap, int nqb,
                        int n_chunks, int rank, int topk_cap, unsigned int* __restrict__ gthr) {
    extern __shared__ __align__(16) char smem[];
    const int tid = threadIdx.x;
    const int q = blockIdx.x;
    const int qb = q / TC_BM, ql = q % TC_BM;
    BlockTopK<FN_THREADS, uint32_t> tk;
    tk.attach(smem, topk_cap, rank, FN_THREADS, BlockTopK<FN_THREADS, uint32_t>::NLO + 3);
    tk.init();
    __syncthreads();
    // every (chunk, query) slot of a sample pass holds exactly TC_SAMPLE_R entries: walk them as one flat list
    const int total = n_chunks * TC_SAMPLE_R;
    for (int base = 0; base < total; base += FN_THREADS) {
        const int e_idx = base + tid;
        const bool valid = e_idx < total;
        unsigned long long e = 0ull;
        if (valid) {
            const int chunk = e_idx / TC_SAMPLE_R, j = e_idx % TC_SAMPLE_R;
            e = cand[(((size_t)(chunk * nqb + qb)) * TC_BM + ql) * cap + j];
        }
        tk.offer(valid, (uint64_t)mono32(__uint_as_float((uint32_t)(e >> 32))), ~(uint32_t)e_idx);
        tk.settle();
    }
    __syncthreads();
    tk.finalize();
    if (tid == 0) gthr[q] = tk.count() >= rank ? (unsigned int)tk.out_hi()[rank - 1] : 0u;
}

// ----------------------------------------------------------------------------------------------- tier-0 gather
// Packs the queries the first pass flagged into a fresh query block matrix (zero rows behind them), publishes how many there
// are (the device-side gate of the re-scan and its finish), and moves the flagged queries beyond the tier's capacity to the
// exact scan's list.  One warp per slot.
__global__ void __launch_bounds__(128)
tier0_gather_kernel(const uint16_t* __restrict__ queries, const int32_t* __restrict__ flag_list, const int32_t* __restrict__ n_flagged,
                    int dim, int n_slots, int max_served, uint16_t* __restrict__ q_out, int32_t* __restrict__ gate,
                    int32_t* __restrict__ flag_list2, unsigned int* __restrict__ gthr) {
    const int lane = threadIdx.x & 31;
    const int slot = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (slot < n_slots && lane == 0) gthr[slot] = 0u;           // the re-scan starts from "no threshold yet"
    const int nf = __ldg(n_flagged);
    const int served = nf < max_served ? nf : max_served;
    if (blockIdx.x == 0 && threadIdx.x == 0) { gate[0] = served; gate[1] = nf - served; }
    if (blockIdx.x == 0)
        for (int i = threadIdx.x; i < nf - served; i += blockDim.x) flag_list2[i] = flag_list[served + i];
    if (slot >= n_slots) return;
    uint4* dst = reinterpret_cast<uint4*>(q_out + (size_t)slot * dim);
    if (slot < served) {
        const uint4* src = reinterpret_cast<const uint4*>(queries + (size_t)__ldg(flag_list + slot) * dim);
        for (int c = lane; c < dim / 8; c += 32) dst[c] = __ldg(src + c);
    } else if (nf > 0) {                                 // (nothing flagged: the re-scan is gated off and never reads the block)
        for (int c = lane; c < dim / 8; c += 32) dst[c] = make_uint4(0u, 0u, 0u, 0u);
    }
}

// ----------------------------------------------------------------------------------------------- host side
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled get_encode() {
    // function-local static: initialised once, thread-safe (C++11)
    static const PFN_tmapEncodeTiled fn = []() -> PFN_tmapEncodeTiled {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            return reinterpret_cast<PFN_tmapEncodeTiled>(ptr);
        return nullptr;
    }();
    return fn;
}

// 2-D tensor map over a row-major [rows, dim] 16-bit matrix; box = 64 elements (one 128-byte swizzle atom) x box_rows.
int make_tensor_map(CUtensorMap* map, const void* base, int64_t rows, int dim, int dtype, int box_rows) {
    PFN_tmapEncodeTiled enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is unavailable (driver too old?)");
        return B200RAG_E_CUDA;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)dim * 2};
    cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, dtype == B200RAG_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                     const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld dim=%d)", (int)r, (long long)rows, dim);
        return B200RAG_E_CUDA;
    }
    return B200RAG_OK;
}

size_t scan3_smem_bytes(int cap, int span);
int scan3_max_clusters(int cap, int span, int sm_count);
int launch_scan3(const void* corpus16, int dtype, const ScanParams& sp, int max_clusters, int span_cap, int span_max, cudaStream_t st);

struct TensorPlan {
    int sm_count, version, cs, tile_rows, nqb, n_tiles, n_chunks, n_items, kprime, cap, topk_cap, max_ctas, qg_span;
    int sample, s_stride, s_tiles, s_chunks, s_items, s_kprime, s_cap, s_rank, s_topk_cap;   // strided sample pass
    int stage_rows;
    size_t scan_smem, finish_smem;
    size_t off_cand, off_cnt, off_gthr, off_flaglist, off_nflag, off_stats, off_qpad, off_exact, off_fin, total;
    // tier 0: flagged queries re-scanned on the tensor path with the widest candidate set the kernels support
    int t0, t0_nq, t0_nqb, t0_kprime, t0_cap, t0_span, t0_chunks, t0_items, t0_clusters;
    int t0_sample, t0_s_stride, t0_s_tiles, t0_s_chunks, t0_s_items;
    size_t t0_smem, off_t0_cand, off_t0_cnt, off_t0_gthr, off_t0_q, off_t0_gate, off_flaglist2;
};

constexpr int TC_T0_MAX_QUERIES = 1024;   // flagged queries one tier-0 pass serves (the rest go straight to the exact scan)
constexpr int TC_T0_KPRIME = 640;         // = TC_MAX_C / 2: tie groups of up to ~500 rows around rank k are resolved

static int gcd_int(int a, int b) { return b ? gcd_int(b, a % b) : a; }

static int sm_count_cached() {
    // one rank drives one GPU model (B200): initialised once, thread-safe (C++11 static)
    static const int n = []() {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
        return v;
    }();
    return n;
}


static TensorPlan plan_tensor(int64_t n_rows, int dim, int n_q, int k) {
    TensorPlan pl;
    pl.sm_count = sm_count_cached();
    pl.nqb = (n_q + TC_BM - 1) / TC_BM;
    pl.kprime = tc_kprime(k);
    pl.cap = tc_bufcap(pl.kprime);
    // kernel generation: 3 = CTA pairs (tcgen05 cta_group::2, dense_tc3.cu) serves two or more query blocks; 1 (single CTA,
    // this file) serves a single block (a pair's M = 256 cannot be filled) and stays selectable for A/B measurements.
    // (A generation 2 -- query block resident in TMEM, TS-mode MMA with N = 64 accumulators, multicast corpus tiles -- was
    // measured at 0.8 PFLOP/s, MMA-issue bound, and removed; it is in the history at commit 1cee701.)
    pl.version = pl.nqb >= 2 ? 3 : 1;
    int forced = option(OPT_SCAN_VERSION, 0);
    if (forced == 1 || forced == 3) pl.version = forced;
    int units;                                              // co-resident scheduling units (CTAs or clusters)
    int qgroups;
    if (pl.version == 3) {
        pl.tile_rows = 256;
        pl.cs = 2;
        qgroups = (pl.nqb + 1) / 2;
        pl.nqb = qgroups * 2;                               // candidate / threshold buffers cover the padded block too
        // tile-major span: as many query-block pairs per work item as fit next to the stage ring in shared memory
        pl.qg_span = qgroups < TC_QG_SPAN_MAX ? qgroups : TC_QG_SPAN_MAX;
        int fs = option(OPT_QG_SPAN, 0);
        if (fs >= 1 && fs <= TC_QG_SPAN_MAX) pl.qg_span = fs < qgroups ? fs : qgroups;
        while (pl.qg_span > 1 && scan3_smem_bytes(pl.cap, pl.qg_span) > TC_SMEM_LIMIT) --pl.qg_span;
        units = scan3_max_clusters(pl.cap, pl.qg_span, pl.sm_count);
        pl.max_ctas = units * 2;
        pl.scan_smem = scan3_smem_bytes(pl.cap, pl.qg_span);
    } else {
        pl.tile_rows = TC_BN;
        pl.cs = 1;
        units = pl.sm_count;
        qgroups = pl.nqb;
        pl.max_ctas = units;
        pl.scan_smem = 1024 + (size_t)TC_STAGES * TC_STAGE_BYTES + (2 * TC_STAGES + 6) * 8 + (size_t)TC_EPI_WARPS * pl.cap * 4;
    }
    pl.n_tiles = (int)((n_rows + pl.tile_rows - 1) / pl.tile_rows);
    const int want = units / gcd_int(qgroups, units);       // smallest chunk count with n_items % units == 0 (span 1)
    // v3 runs tile-major: one work item covers a span of query-block pairs for every tile of its chunk
    if (pl.version != 3) pl.qg_span = 1;
    const int n_spans = (qgroups + pl.qg_span - 1) / pl.qg_span;
    const int want_main = units / gcd_int(n_spans, units);
    pl.n_chunks = want_main < pl.n_tiles ? want_main : pl.n_tiles;
    if (pl.n_chunks < 1) pl.n_chunks = 1;
    pl.n_items = n_spans * pl.n_chunks;
    // strided SAMPLE pass that seeds the per-query thresholds: every s_stride-th tile is scored, the epilogue keeps the
    // TC_SAMPLE_R best 32-row group maxima per (chunk, query) in registers, and the s_rank-th best over all chunks becomes
    // the query's starting threshold.  About s_rank * s_stride rows of the whole corpus beat it; that product is held
    // near 8 k' (see below).
    // Sizing.  The threshold is the r-th best of a 1/stride sample, so the number of corpus rows above it is about
    // stride * Gamma(r): with r = 16 and r * stride = 8 k' the chance that fewer than k' rows survive (the query is then
    // flagged and re-run on the slow exact path) is P(Gamma(16) < 2) ~ 5e-10 per query; at r = 8, 6 k' it was 6e-5 and
    // showed up as a handful of fallbacks per thousand batches.  12 k' (1e-12) cost 5 % of the scan at 1.25M-row shards:
    // every survivor is appended by the epilogue and gathered again by the finish kernel.
    pl.s_rank = TC_SAMPLE_R;
    pl.s_stride = option(OPT_SAMPLE_MULT, 8) * pl.kprime / TC_SAMPLE_R;       // (A/B knob: rows above the threshold, in k')
    if (pl.s_stride < 1) pl.s_stride = 1;
    {
        // Small shards searched deeply (125K rows, k = 500: 8 k' would be a single sampled tile).  The sample ranks 32-row GROUP
        // maxima: with G sampled groups the r-th best of them sits where a group's maximum exceeds it with probability
        // r / (G + 1), i.e. a row does with probability -ln(1 - r / (G + 1)) / 32  (= r / (32 G), the r * stride rule, while
        // G >> r; more rows than that rule says when G is a few r, down to ~9 % of the corpus at G = r).  Sampled tiles are
        // given up one at a time, never below r groups, until the estimate reaches 6 k'; the pass runs if it reaches 5 k'
        // (chance of fewer than k' survivors, which costs the flagged re-scan: < 1e-6 per query) and is skipped otherwise.
        const int groups_per_tile = pl.tile_rows / 32;
        const int s_min_tiles = (TC_SAMPLE_R + groups_per_tile - 1) / groups_per_tile;
        auto rows_above = [&](int tiles) {
            const double frac = (double)pl.s_rank / ((double)tiles * groups_per_tile + 1.0);
            return frac < 1.0 ? (double)n_rows * -log1p(-frac) / 32.0 : 0.0;
        };
        int tiles = pl.n_tiles / pl.s_stride;
        if (tiles < 4 * s_min_tiles) tiles = 4 * s_min_tiles < pl.n_tiles ? 4 * s_min_tiles : pl.n_tiles;
        while (tiles > s_min_tiles && rows_above(tiles) < 6.0 * pl.kprime) --tiles;
        pl.s_stride = tiles > 0 ? pl.n_tiles / tiles : 1;
        if (pl.s_stride < 1) pl.s_stride = 1;
        pl.s_tiles = pl.n_tiles / pl.s_stride;
        pl.sample = pl.s_tiles >= s_min_tiles && rows_above(pl.s_tiles) >= 5.0 * pl.kprime && option(OPT_NO_SAMPLE, 0) == 0;
    }
    pl.s_chunks = want < pl.s_tiles ? want : (pl.s_tiles > 0 ? pl.s_tiles : 1);
    pl.s_items = qgroups * pl.s_chunks;
    pl.s_kprime = TC_SAMPLE_R;
    pl.s_cap = TC_SAMPLE_R;
    pl.s_topk_cap = BlockTopK<FN_THREADS, uint32_t>::capacity_for(pl.s_rank, FN_THREADS);
    pl.topk_cap = BlockTopK<FN_THREADS, uint32_t>::capacity_for(pl.kprime, FN_THREADS);
    pl.finish_smem = (size_t)dim * 8 + (size_t)pl.kprime * (8 + 4) + (size_t)pl.n_chunks * 4 + 32 +
                     BlockTopK<FN_THREADS, uint32_t>::smem_bytes(pl.topk_cap) + 64;
    pl.stage_rows = FN_THREADS / 8;                             // staged candidate rows of one re-score batch
    { int sr = option(OPT_STAGE_ROWS, 0); if (sr == 8 || sr == 16 || sr == 32) pl.stage_rows = sr; }
    while (pl.stage_rows > 1 && pl.finish_smem + (size_t)pl.stage_rows * ((size_t)dim * 2 + 16) > 160 * 1024) pl.stage_rows /= 2;
    pl.finish_smem += (size_t)pl.stage_rows * ((size_t)dim * 2 + 16);
    size_t off = 0;
    auto take = [&](size_t bytes) { off = align_up(off, 256); size_t o = off; off += bytes; return o; };
    const int max_chunks = pl.n_chunks > pl.s_chunks ? pl.n_chunks : pl.s_chunks;
    pl.off_cand = take((size_t)max_chunks * pl.nqb * TC_BM * (pl.cap > pl.s_cap ? pl.cap : pl.s_cap) * 8);
    pl.off_cnt = take((size_t)max_chunks * pl.nqb * TC_BM * 4);
    pl.off_gthr = take((size_t)pl.nqb * TC_BM * 4);
    pl.off_flaglist = take((size_t)n_q * 4);
    pl.off_qpad = take((size_t)pl.nqb * TC_BM * dim * 2);      // query block padded with zero rows (no TMA out-of-bounds rows)
    pl.off_nflag = take(256);
    pl.off_fin = take(finish3_workspace_bytes(n_q, pl.kprime));          // split finish: selections + exact scores
    // tier 0 (CTA-pair scan generation only; pointless when the first pass already runs at the widest k')
    pl.t0 = option(OPT_NO_TIER0, 0) == 0 && pl.kprime < TC_T0_KPRIME && n_rows > 0;
    if (pl.t0) {
        pl.t0_nq = n_q < TC_T0_MAX_QUERIES ? n_q : TC_T0_MAX_QUERIES;
        pl.t0_nqb = 2 * ((pl.t0_nq + 2 * TC_BM - 1) / (2 * TC_BM));
        pl.t0_kprime = TC_T0_KPRIME;
        pl.t0_cap = tc_bufcap(pl.t0_kprime);
        const int qg = pl.t0_nqb / 2;
        pl.t0_span = qg < TC_QG_SPAN_MAX ? qg : TC_QG_SPAN_MAX;
        while (pl.t0_span > 1 && scan3_smem_bytes(pl.t0_cap, pl.t0_span) > TC_SMEM_LIMIT) --pl.t0_span;
        pl.t0_smem = scan3_smem_bytes(pl.t0_cap, pl.t0_span);
        pl.t0_clusters = scan3_max_clusters(pl.t0_cap, pl.t0_span, pl.sm_count);
        const int spans = (qg + pl.t0_span - 1) / pl.t0_span;
        const int want_t0 = pl.t0_clusters / gcd_int(spans, pl.t0_clusters);
        const int tiles3 = (int)((n_rows + 255) / 256);
        pl.t0_chunks = want_t0 < tiles3 ? want_t0 : tiles3;
        if (pl.t0_chunks < 1) pl.t0_chunks = 1;
        pl.t0_items = spans * pl.t0_chunks;
        // strided sample pass of the re-scan.  About 4 k' rows above the threshold instead of the first pass's 8 k': at
        // k' = 640 every survivor costs the epilogue's slow path and the finish kernel's gather (8 k': 1.64 ms for the re-scan
        // of 1M rows, B = 1024), and a query that ends up with fewer than k' survivors here -- P(Gamma(16) < 4) ~ 4e-6 -- is
        // still answered exactly, by the last tier.
        pl.t0_s_stride = TC_T0_SAMPLE_MULT * pl.t0_kprime / TC_SAMPLE_R;
        if (pl.t0_s_stride > tiles3 / 4) pl.t0_s_stride = tiles3 / 4;
        if (pl.t0_s_stride < 1) pl.t0_s_stride = 1;
        pl.t0_s_tiles = tiles3 / pl.t0_s_stride;
        pl.t0_sample = pl.t0_s_tiles >= 4 && TC_SAMPLE_R * pl.t0_s_stride >= 3 * pl.t0_kprime && option(OPT_NO_SAMPLE, 0) == 0;
        const int want_s = pl.t0_clusters / gcd_int(qg, pl.t0_clusters);
        pl.t0_s_chunks = want_s < pl.t0_s_tiles ? want_s : (pl.t0_s_tiles > 0 ? pl.t0_s_tiles : 1);
        pl.t0_s_items = qg * pl.t0_s_chunks;
        if (pl.t0_smem > TC_SMEM_LIMIT || finish2_smem_bytes(dim, pl.t0_kprime) > 200 * 1024) pl.t0 = 0;
    }
    if (pl.t0) {
        const int t0_max_chunks = pl.t0_chunks > pl.t0_s_chunks ? pl.t0_chunks : pl.t0_s_chunks;
        pl.off_t0_cand = take((size_t)t0_max_chunks * pl.t0_nqb * TC_BM * pl.t0_cap * 8);
        pl.off_t0_cnt = take((size_t)t0_max_chunks * pl.t0_nqb * TC_BM * 4);
        pl.off_t0_gthr = take((size_t)pl.t0_nqb * TC_BM * 4);
        pl.off_t0_q = take((size_t)pl.t0_nqb * TC_BM * dim * 2);
        pl.off_t0_gate = take(256);                          // [0] live tier-0 slots, [1] queries left for the exact scan
        pl.off_flaglist2 = take((size_t)n_q * 4);
    }
    pl.off_stats = take((size_t)256 * ST_N * 8);
    {   // exact fallback, tier A (<= TC_FALLBACK_BATCH queries) and tier B (the rest): they run one after the other in the same region
        const int na = n_q < TC_FALLBACK_BATCH ? n_q : TC_FALLBACK_BATCH;
        size_t ea = exact_workspace_bytes(n_rows, dim, na, k);
        size_t eb = n_q > na ? exact_workspace_bytes(n_rows, dim, n_q - na, k) : 0;
        pl.off_exact = take(ea > eb ? ea : eb);
    }
    pl.total = align_up(off, 256);
    return pl;
}

size_t tensor_workspace_bytes(int64_t n_rows, int dim, int n_q, int k) { return plan_tensor(n_rows, dim, n_q, k).total; }

// Can the tensor-core path serve this shape?  (k beyond the candidate-buffer limit, or more rows than 32-bit local row
// numbers address, go to the exact CUDA-core scan; AUTO mode does that on its own.)
bool tensor_supported(int64_t n_rows, int dim, int n_q, int k) {
    const TensorPlan pl = plan_tensor(n_rows, dim, n_q, k);
    return !(pl.cap > TC_MAX_C || pl.scan_smem > TC_SMEM_LIMIT || n_rows >= ((int64_t)1 << 32) - TC_BN);
}


static thread_local cudaEvent_t g_prof_start = nullptr, g_prof_stop = nullptr;
void profile_next_scan(void* a, void* b) { g_prof_start = static_cast<cudaEvent_t>(a); g_prof_stop = static_cast<cudaEvent_t>(b); }

int run_tensor(const void* corpus16, int64_t n_rows, int dim, int dtype, const void* queries16, int n_q, int k,
               int64_t id_offset, double* out_scores, int64_t* out_ids, int32_t* out_flags, double row_norm_bound,
               float* out_err, int with_fallback, void* workspace, size_t workspace_bytes, cudaStream_t st,
               const uint32_t* row_mask) {
    const int approx = with_fallback == 2;       // B200RAG_DENSE_APPROX
    if (approx) with_fallback = 0;
    TensorPlan pl = plan_tensor(n_rows, dim, n_q, k);
    if (pl.cap > TC_MAX_C || pl.scan_smem > TC_SMEM_LIMIT || n_rows >= ((int64_t)1 << 32) - TC_BN) {
        set_error("dense_topk(tensor): k=%d (k'=%d) or n_rows=%lld beyond the tensor-core path limits; use B200RAG_DENSE_EXACT",
                  k, pl.kprime, (long long)n_rows);
        return B200RAG_E_UNSUPPORTED;
    }
    if (workspace_bytes < pl.total) {
        set_error("dense_topk(tensor): workspace too small (%zu < %zu)", workspace_bytes, pl.total);
        return B200RAG_E_WORKSPACE;
    }
    char* ws = static_cast<char*>(workspace);
    int32_t* flag_list = reinterpret_cast<int32_t*>(ws + pl.off_flaglist);
    int32_t* n_flagged = reinterpret_cast<int32_t*>(ws + pl.off_nflag);
    B200_CUDA_CHECK(cudaMemsetAsync(n_flagged, 0, sizeof(int32_t), st));
    B200_CUDA_CHECK(cudaMemsetAsync(ws + pl.off_gthr, 0, (size_t)pl.nqb * TC_BM * 4, st));

    if (n_rows > 0) {
        ScanParams sp;
        sp.n_rows = n_rows;
        sp.n_q = n_q;
        sp.dim = dim;
        sp.n_kblocks = (dim + TC_BK - 1) / TC_BK;
        sp.nqb = pl.nqb;
        sp.idesc = pl.version == 3 ? umma_idesc_mn(dtype, 2 * TC_BM, pl.tile_rows) : umma_idesc(dtype, pl.tile_rows);
        sp.cand = reinterpret_cast<unsigned long long*>(ws + pl.off_cand);
        sp.cand_cnt = reinterpret_cast<int*>(ws + pl.off_cnt);
        sp.gthr = reinterpret_cast<unsigned int*>(ws + pl.off_gthr);
        // The scan reads whole 128-query blocks.  A partially out-of-bounds TMA box is legal but measurably slow (batch 1:
        // 3.5 ms vs 2.4 ms for the same 15.4 GB), so a ragged batch is copied once into a zero-padded block buffer.
        const void* q_scan = queries16;
        int n_q_scan = n_q;
        if (n_q != pl.nqb * TC_BM) {
            char* qpad = ws + pl.off_qpad;
            const size_t used = (size_t)n_q * dim * 2, total_q = (size_t)pl.nqb * TC_BM * dim * 2;
            B200_CUDA_CHECK(cudaMemcpyAsync(qpad, queries16, used, cudaMemcpyDeviceToDevice, st));
            B200_CUDA_CHECK(cudaMemsetAsync(qpad + used, 0, total_q - used, st));
            q_scan = qpad;
            n_q_scan = pl.nqb * TC_BM;
        }
        sp.queries = static_cast<const uint16_t*>(q_scan);
        sp.row_mask = row_mask;
        sp.gate = nullptr;
        // per-role cycle counters go to the caller's buffer of this thread (b200rag_debug_set_stats_buffer), if any
        sp.stats = stats_buffer(STATS_SCAN, (size_t)256 * ST_N);
        if (sp.stats) B200_CUDA_CHECK(cudaMemsetAsync(sp.stats, 0, (size_t)256 * ST_N * 8, st));
        CUtensorMap map_q, map_x;
        if (pl.version == 1) {
            int rc = make_tensor_map(&map_q, q_scan, n_q_scan, dim, dtype, TC_BM);   // whole blocks
            if (rc) return rc;
            rc = make_tensor_map(&map_x, corpus16, n_rows, dim, dtype, TC_BN);
            if (rc) return rc;
            B200_CUDA_CHECK(cudaFuncSetAttribute(dense_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.scan_smem));
        }
        auto launch = [&](const ScanParams& spx) -> int {
            if (pl.version == 3) return launch_scan3(corpus16, dtype, spx, pl.max_ctas / 2, pl.cap, pl.qg_span, st);
            int grid = spx.n_items < pl.sm_count ? spx.n_items : pl.sm_count;
            dense_scan_kernel<<<grid, TC_THREADS, pl.scan_smem, st>>>(map_q, map_x, spx); count_launch();
            B200_CUDA_CHECK(cudaGetLastError());
            return B200RAG_OK;
        };
        const bool prof = g_prof_start && g_prof_stop;
        sp.sample = 0;
        sp.qg_span = 1;
        if (pl.sample) {
            // pass 0: every s_stride-th tile, tiny k'; its r-th best score per query seeds the thresholds
            ScanParams s0 = sp;
            s0.stats = nullptr;
            s0.n_tiles = pl.s_tiles;
            s0.tile_stride = pl.s_stride;
            s0.n_chunks = pl.s_chunks;
            s0.n_items = pl.s_items;
            s0.kprime = pl.s_kprime;
            s0.cap = pl.s_cap;
            s0.sample = 1;
            s0.qg_span = 1;
            int rc = launch(s0);
            if (rc) return rc;
            if (option(OPT_FINISH_VERSION, 0) == 1) {
                size_t tsm = BlockTopK<FN_THREADS, uint32_t>::smem_bytes(pl.s_topk_cap) + 64;
                sample_threshold_kernel<<<pl.nqb * TC_BM, FN_THREADS, tsm, st>>>(s0.cand, s0.cand_cnt, s0.cap, pl.nqb, pl.s_chunks,
                                                                                pl.s_rank, pl.s_topk_cap, sp.gthr); count_launch();
                B200_CUDA_CHECK(cudaGetLastError());
            } else {
                rc = launch_sample_threshold2(s0.cand, s0.cap, pl.nqb, pl.s_chunks, pl.s_rank, sp.gthr, st, nullptr);
                if (rc) return rc;
            }
        }
        sp.n_tiles = pl.n_tiles;
        sp.tile_stride = 1;
        sp.n_chunks = pl.n_chunks;
        sp.n_items = pl.n_items;
        sp.kprime = pl.kprime;
        sp.cap = pl.cap;
        sp.qg_span = pl.qg_span;
        if (prof) B200_CUDA_CHECK(cudaEventRecord(g_prof_start, st));     // the hook brackets the FULL scan kernel only
        int rc = launch(sp);
        if (rc) return rc;
        if (prof) {
            B200_CUDA_CHECK(cudaEventRecord(g_prof_stop, st));
            g_prof_start = g_prof_stop = nullptr;
        }
    } else {
        B200_CUDA_CHECK(cudaMemsetAsync(ws + pl.off_cnt, 0, (size_t)pl.n_chunks * pl.nqb * TC_BM * 4, st));
    }

    FinishParams fp;
    fp.corpus = static_cast<const uint16_t*>(corpus16);
    fp.queries = static_cast<const uint16_t*>(queries16);
    fp.n_rows = n_rows;
    fp.dim = dim;
    fp.n_q = n_q;
    fp.k = k;
    fp.kprime = pl.kprime;
    fp.cap = pl.cap;
    fp.nqb = pl.nqb;
    fp.n_chunks = n_rows > 0 ? pl.n_chunks : 0;
    fp.topk_cap = pl.topk_cap;
    fp.stage_rows = pl.stage_rows;
    fp.id_offset = id_offset;
    fp.row_norm_bound = row_norm_bound;
    fp.cand = reinterpret_cast<const unsigned long long*>(ws + pl.off_cand);
    fp.cand_cnt = reinterpret_cast<const int*>(ws + pl.off_cnt);
    fp.gthr = reinterpret_cast<const unsigned int*>(ws + pl.off_gthr);
    fp.out_scores = out_scores;
    fp.out_ids = out_ids;
    fp.out_flags = out_flags;
    fp.flag_list = flag_list;
    fp.n_flagged = n_flagged;
    fp.err_max = out_err;
    fp.approx = approx;
    // Which finish kernel: the second generation (dense_finish.cu: 128 threads, warp-level selection) wins while k' is small --
    // 118 vs 183 us at B = 1024, k = 100 -- but its selection is one warp's serial walk over ~8 k' survivors and its rank
    // all-pairs, so the first generation (256 threads, block-wide streaming top-k, bitonic sort) is faster for deep searches:
    // B = 256, k = 500: 0.80 vs 1.19 ms per search, k = 200: 0.60 vs 0.69 ms.  Option finish_version: 0 auto, 1 / 2 forced.
    const int fin = option(OPT_FINISH_VERSION, 0);
    const bool fin2 = fin == 2 || (fin == 0 && pl.kprime <= 160);
    if (!approx && fin == 3 && finish2_smem_bytes(dim, pl.kprime) <= 200 * 1024 && n_rows > 0) {
        int rc = launch_finish3(fp, dtype, ws + pl.off_fin, st);
        if (rc) return rc;
    } else if (approx || (fin2 && finish2_smem_bytes(dim, pl.kprime) <= 200 * 1024)) {
        int rc = launch_finish2(fp, dtype, st);
        if (rc) return rc;
    } else if (dtype == B200RAG_F16) {
        B200_CUDA_CHECK(cudaFuncSetAttribute(dense_finish_kernel<B200RAG_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.finish_smem));
        dense_finish_kernel<B200RAG_F16><<<n_q, FN_THREADS, pl.finish_smem, st>>>(fp); count_launch();
    } else {
        B200_CUDA_CHECK(cudaFuncSetAttribute(dense_finish_kernel<B200RAG_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.finish_smem));
        dense_finish_kernel<B200RAG_BF16><<<n_q, FN_THREADS, pl.finish_smem, st>>>(fp); count_launch();
    }
    B200_CUDA_CHECK(cudaGetLastError());

    const int32_t* fb_list = flag_list;          // what the exact scan has to redo: all flagged queries, unless tier 0 runs first
    const int32_t* fb_count = n_flagged;
    if (with_fallback && pl.t0) {
        // Tier 0: the flagged queries once more on the tensor path, with k' = 640 candidates instead of k + 28.  A query is
        // flagged when rows tie (to within the fp32 error band) across the edge of its candidate set -- duplicated chunks,
        // boilerplate -- and a tie group of up to ~500 rows fits the wider set, so the proof then succeeds and the ~100x
        // slower CUDA-core scan is not needed.  Gated on the device by the flagged-query count: no host round trip, and a
        // batch without flagged queries pays three empty launches.
        uint16_t* q_t0 = reinterpret_cast<uint16_t*>(ws + pl.off_t0_q);
        int32_t* gate = reinterpret_cast<int32_t*>(ws + pl.off_t0_gate);
        int32_t* flag_list2 = reinterpret_cast<int32_t*>(ws + pl.off_flaglist2);
        const int n_slots = pl.t0_nqb * TC_BM;
        tier0_gather_kernel<<<(n_slots + 3) / 4, 128, 0, st>>>(static_cast<const uint16_t*>(queries16), flag_list, n_flagged, dim, n_slots,
                                                               pl.t0_nq, q_t0, gate, flag_list2,
                                                               reinterpret_cast<unsigned int*>(ws + pl.off_t0_gthr)); count_launch();
        B200_CUDA_CHECK(cudaGetLastError());
        ScanParams st0;
        st0.n_rows = n_rows;
        st0.n_q = n_slots;
        st0.dim = dim;
        st0.n_kblocks = (dim + TC_BK - 1) / TC_BK;
        st0.n_tiles = (int)((n_rows + 255) / 256);
        st0.tile_stride = 1;
        st0.nqb = pl.t0_nqb;
        st0.n_chunks = pl.t0_chunks;
        st0.n_items = pl.t0_items;
        st0.qg_span = pl.t0_span;
        st0.kprime = pl.t0_kprime;
        st0.cap = pl.t0_cap;
        st0.sample = 0;
        st0.idesc = umma_idesc_mn(dtype, 2 * TC_BM, 256);
        st0.cand = reinterpret_cast<unsigned long long*>(ws + pl.off_t0_cand);
        st0.cand_cnt = reinterpret_cast<int*>(ws + pl.off_t0_cnt);
        st0.gthr = reinterpret_cast<unsigned int*>(ws + pl.off_t0_gthr);
        st0.queries = q_t0;
        st0.row_mask = row_mask;
        st0.stats = nullptr;
        st0.gate = gate;
        int rc;
        if (pl.t0_sample) {
            ScanParams ss = st0;
            ss.n_tiles = pl.t0_s_tiles;
            ss.tile_stride = pl.t0_s_stride;
            ss.n_chunks = pl.t0_s_chunks;
            ss.n_items = pl.t0_s_items;
            ss.qg_span = 1;
            ss.kprime = TC_SAMPLE_R;
            ss.cap = TC_SAMPLE_R;
            ss.sample = 1;
            rc = launch_scan3(corpus16, dtype, ss, pl.t0_clusters, pl.t0_cap, pl.t0_span, st);
            if (rc) return rc;
            rc = launch_sample_threshold2(ss.cand, ss.cap, pl.t0_nqb, pl.t0_s_chunks, TC_SAMPLE_R, st0.gthr, st, gate);
            if (rc) return rc;
        }
        rc = launch_scan3(corpus16, dtype, st0, pl.t0_clusters, pl.t0_cap, pl.t0_span, st);
        if (rc) return rc;
        FinishParams f0 = fp;
        f0.n_q = n_slots;
        f0.kprime = pl.t0_kprime;
        f0.cap = pl.t0_cap;
        f0.nqb = pl.t0_nqb;
        f0.n_chunks = pl.t0_chunks;
        f0.cand = st0.cand;
        f0.cand_cnt = st0.cand_cnt;
        f0.gthr = st0.gthr;
        f0.flag_list = flag_list2;
        f0.n_flagged = gate + 1;
        f0.q_list = flag_list;
        f0.gate = gate;
        rc = launch_finish2(f0, dtype, st);
        if (rc) return rc;
        fb_list = flag_list2;
        fb_count = gate + 1;
    }
    if (with_fallback) {
        // AUTO: results must be exact for every query, so the queries that are still flagged are re-run on the exact path --
        // without a host round trip: the exact scan is launched unconditionally over the flag list and gated ON THE DEVICE by the
        // flagged-query counter (CTAs of unused slots leave at once).  Tier A covers the first TC_FALLBACK_BATCH flagged
        // queries with a many-chunk plan (the usual case: zero, one or a few of them); tier B covers every further slot
        // with the few-chunk plan of a large batch and only does work on pathological inputs (massive ties).
        const int na = n_q < TC_FALLBACK_BATCH ? n_q : TC_FALLBACK_BATCH;
        int rc = run_exact(corpus16, n_rows, dim, dtype, queries16, na, fb_list, k, id_offset, out_scores, out_ids,
                           ws + pl.off_exact, pl.total - pl.off_exact, st, fb_count, 0, row_mask);
        if (rc) return rc;
        if (n_q > na) {
            rc = run_exact(corpus16, n_rows, dim, dtype, queries16, n_q - na, fb_list + na, k, id_offset, out_scores, out_ids,
                           ws + pl.off_exact, pl.total - pl.off_exact, st, fb_count, na, row_mask);
            if (rc) return rc;
        }
    }
    return B200RAG_OK;
}

}  // namespace b200rag
