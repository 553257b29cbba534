// dense_finish.cu -- finish stage of the tensor-core dense scan, second generation  (K2).
//
// Per query: merge the per-chunk candidate lists the scan epilogue wrote to the k' best by tensor-core score, RE-SCORE them in
// the canonical fp64 arithmetic (bit-identical to oracle/exact_scan.c), rank by (score desc, row asc), emit the top k and
// PROVE completeness (see dense_tc.cu's header).  Round 1 ran one 256-thread CTA per query around a block-level streaming
// top-k: 63 M warp instructions for 1024 queries, three quarters of them in the selection's barriers, histograms and
// bitonic passes (profiles/r1: 0.18 ms, 37 % issue utilisation, 3 CTAs per SM).  This generation:
//   * one 128-thread CTA per query, ~40 KB of shared memory -> 5 CTAs per SM, no block-level selection at all;
//   * warp 0 streams the chunk lists -- 32 survivors at a time, across as many chunks as they span -- through a 512-entry buffer
//     and keeps the k' greatest with the same register-resident radix select the scan epilogue uses (tc_common.cuh:
//     warp_compact): typically two or three compactions per query; meanwhile warps 1..3 convert the query to fp64;
//   * re-score from rows staged with warp-wide 16-byte cp.async copies (coalesced 512-byte requests), eight threads per row,
//     one per canonical fp64 lane.  (A first version of this generation let one thread walk a whole row straight from global
//     memory: 16-byte requests from 32 different rows per warp instruction -- 262 us, slower than round 1.)
//   * all-pairs rank count in shared memory (k' <= 640 -> at most 3200 compares per thread), no sort.
// Which candidates survive a tie at the k'-th tensor-core score is immaterial: every dropped row has a tensor-core score
// <= m either way, so the proof -- and with it the exact result -- does not depend on it.
#include "finish.cuh"

namespace b200rag {

constexpr int F2_THREADS = 128;
constexpr int F2_ROWS = F2_THREADS / 8;     // candidate rows re-scored per batch: eight threads share a row
constexpr int F2_GATHER = 1024;             // survivors gathered per round (8 KB, aliased with the row staging buffer)

__host__ __device__ inline int finish2_sel_cap(int kprime) { return 2 * kprime <= 512 ? 512 : 2 * kprime; }
constexpr int F2_SEG = 384;                 // elements of a row staged at a time (768 bytes): the staging buffer does not grow with dim,
                                            // so eight CTAs fit an SM and 1024 queries finish in ONE wave (at 16 whole 768-d rows:
                                            // five per SM, 1.4 waves)
__host__ __device__ inline int finish2_seg(int dim) { return dim < F2_SEG ? dim : F2_SEG; }

size_t finish2_smem_bytes(int dim, int kprime) {
    const int sel = finish2_sel_cap(kprime);
    size_t stage = (size_t)F2_ROWS * ((size_t)finish2_seg(dim) * 2 + 16);
    if (stage < (size_t)F2_GATHER * 8) stage = (size_t)F2_GATHER * 8;
    return (size_t)dim * 8 + (size_t)sel * 8 + (size_t)sel * 4 + (size_t)kprime * (8 + 4) + 256 * 4 + stage + 64;
}

template <int DTYPE>
__global__ void __launch_bounds__(F2_THREADS) dense_finish2_kernel(const FinishParams p) {
    extern __shared__ __align__(16) char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // slot = position of the query in THIS launch's query block matrix (candidate buffers, thresholds); q = the query the
    // results belong to.  They differ in the tier-0 re-scan, which packs the flagged queries of the first pass.
    const int slot_q = blockIdx.x;
    if (p.gate && slot_q >= __ldg(p.gate)) return;
    const int q = p.q_list ? __ldg(p.q_list + slot_q) : slot_q;
    const int qb = slot_q / TC_BM, ql = slot_q % TC_BM;
    const int sel_cap = finish2_sel_cap(p.kprime);
    double* qd = reinterpret_cast<double*>(smem);                                         // [dim] the query in fp64
    unsigned long long* buf = reinterpret_cast<unsigned long long*>(qd + p.dim);          // [sel_cap] (score bits << 32) | row
    uint32_t* scratch = reinterpret_cast<uint32_t*>(buf + sel_cap);                        // [sel_cap] keys of the smem select
    double* exact = reinterpret_cast<double*>(scratch + sel_cap);                          // [kprime]
    uint32_t* rows = reinterpret_cast<uint32_t*>(exact + p.kprime);                        // [kprime]
    int* s_cnt = reinterpret_cast<int*>(rows + p.kprime);                                  // [<= 256] survivors per chunk
    char* stage = reinterpret_cast<char*>(s_cnt + 256);
    stage = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(stage) + 15) & ~uintptr_t(15));
    __shared__ int s_n, s_total;
    __shared__ float s_m;
    __shared__ double s_q2[F2_THREADS / 32];
    __shared__ float s_err;
    __shared__ double s_ek_sh;

    // how many survivors did the scan leave in every chunk of this query?  (all threads, one round trip)
    for (int c = tid; c < p.n_chunks; c += F2_THREADS) s_cnt[c] = __ldg(p.cand_cnt + ((size_t)(c * p.nqb + qb)) * TC_BM + ql);
    __syncthreads();
    if (warp == 0) {                                     // exclusive prefix of the counts, in place (<= 256 chunks)
        int carry = 0;
        for (int c0 = 0; c0 < p.n_chunks; c0 += 32) {
            const int c = c0 + lane;
            const int v = c < p.n_chunks ? s_cnt[c] : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += u;
            }
            if (c < p.n_chunks) s_cnt[c] = carry + incl - v;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) s_total = carry;
    }
    __syncthreads();
    const int total = s_total;
    // ---- 1. the k' best by tensor-core score over all chunks of this query.  All threads first copy the survivors -- a handful
    //         per chunk, ~8 k' in all -- into shared memory with independent loads (one round trip instead of one per 32
    //         entries), F2_GATHER at a time into the region the re-score later stages rows in; warp 0 then selects from there.
    unsigned long long* gathered = reinterpret_cast<unsigned long long*>(stage);
    int cnt = 0;                                         // (warp 0's running state across the rounds)
    float thr = -CUDART_INF_F;
    bool compacted = false;
    for (int g0 = 0; g0 < total; g0 += F2_GATHER) {
        const int g_n = min(F2_GATHER, total - g0);
        for (int i = tid; i < g_n; i += F2_THREADS) {
            const int pos = g0 + i;
            int lo = 0, hi = p.n_chunks - 1;             // last chunk whose prefix is <= pos
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (s_cnt[mid] <= pos) lo = mid;
                else hi = mid - 1;
            }
            gathered[i] = __ldg(p.cand + (((size_t)(lo * p.nqb + qb)) * TC_BM + ql) * p.cap + (pos - s_cnt[lo]));
        }
        __syncthreads();
        if (warp == 0) {
            for (int i0 = 0; i0 < g_n; i0 += 32) {
                const int i = i0 + lane;
                unsigned long long e = 0ull;
                bool pass = false;
                if (i < g_n) {
                    e = gathered[i];
                    pass = !compacted || __uint_as_float((uint32_t)(e >> 32)) > thr;
                }
                const unsigned bal = __ballot_sync(0xffffffffu, pass);
                if (pass) buf[cnt + __popc(bal & ((1u << lane) - 1u))] = e;
                cnt += __popc(bal);
                if (cnt > sel_cap - 32) {
                    __syncwarp();
                    thr = fmaxf(thr, warp_compact(buf, cnt, p.kprime, sel_cap, smem_u32(scratch), lane));
                    cnt = p.kprime;
                    compacted = true;
                }
            }
        }
        __syncthreads();
    }
    if (warp == 0) {
        __syncwarp();
        if (cnt > p.kprime) {
            thr = fmaxf(thr, warp_compact(buf, cnt, p.kprime, sel_cap, smem_u32(scratch), lane));
            cnt = p.kprime;
            compacted = true;
        }
        if (lane == 0) {
            s_n = cnt;
            // m: every row outside the candidate set has tensor-core score <= m (-inf if nothing was ever dropped): the scan
            // drops a row only below a threshold it pushed to gthr, the merge above only at or below its last threshold
            const unsigned int gk = p.gthr[slot_q];
            s_m = fmaxf(compacted ? thr : -CUDART_INF_F, gk ? unmono32(gk) : -CUDART_INF_F);
            s_err = 0.f;
            s_ek_sh = -CUDART_INF;
        }
    } else if (!p.fin_sel) {
        // ---- the query in fp64 and its squared norm
        double q2 = 0.0;
        for (int d = tid - 32; d < p.dim; d += F2_THREADS - 32) {
            const double v = bits_to_double<DTYPE>(p.queries[(size_t)q * p.dim + d]);
            qd[d] = v;
            q2 = fma(v, v, q2);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) q2 += __shfl_xor_sync(0xffffffffu, q2, off);
        if (lane == 0) s_q2[warp] = q2;
    }
    __syncthreads();
    const int n = s_n;
    if (p.fin_sel) {                                     // split finish: hand the selection over and leave
        for (int i = tid; i < n; i += F2_THREADS) p.fin_sel[(size_t)slot_q * p.kprime + i] = buf[i];
        if (tid == 0) {
            p.fin_n[slot_q] = n;
            p.fin_m[slot_q] = s_m;
            if (p.err_max) p.err_max[q] = 0.f;
        }
        return;
    }

    // ---- 2. exact canonical re-score.  Rows are staged through shared memory F2_ROWS at a time with 16-byte cp.async copies
    //         (a warp copies whole rows: coalesced 512-byte requests, nothing passes through registers), then eight threads
    //         share a row, one per canonical lane (lane j sums d = j mod 8 in increasing d, exactly as oracle/exact_scan.c
    //         does), and the lanes are combined in the canonical tree with shuffles.
    float my_err = 0.f;
    if (p.approx) {
        // approximate mode: the candidates' tensor-core scores ARE the result scores
        for (int i = tid; i < n; i += F2_THREADS) {
            const unsigned long long e = buf[i];
            rows[i] = (uint32_t)e;
            exact[i] = (double)__uint_as_float((uint32_t)(e >> 32));
        }
    } else {
        constexpr int RB = F2_ROWS;
        const int l8 = tid & 7, grp = tid >> 3;
        const int seg = finish2_seg(p.dim);
        const int stride = seg * 2 + 16;                          // staged row pitch: +16 bytes keeps the rows of a warp on different banks
        const double* qj = qd + l8;
        for (int i0 = 0; i0 < n; i0 += RB) {
            const int rows_here = min(RB, n - i0);
            const int i = i0 + grp;
            const bool valid = i < n;
            const uint16_t* x = reinterpret_cast<const uint16_t*>(stage + (size_t)grp * stride) + l8;
            double acc = 0.0;
            for (int d0 = 0; d0 < p.dim; d0 += seg) {             // (a lane's chain runs on across the segments, in increasing d)
                const int dl = min(seg, p.dim - d0);
                for (int r = warp; r < rows_here; r += F2_THREADS / 32) {
                    const uint32_t row = (uint32_t)buf[i0 + r];
                    const uint4* src = reinterpret_cast<const uint4*>(p.corpus + (size_t)row * p.dim + d0);
                    const uint32_t dst = smem_u32(stage + (size_t)r * stride);
                    for (int c = lane; c < dl / 8; c += 32)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + c * 16), "l"(src + c) : "memory");
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncthreads();
                if (valid) {
#pragma unroll 4
                    for (int d = 0; d < dl; d += 8) acc = fma(qj[d0 + d], bits_to_double<DTYPE>(x[d]), acc);
                }
                __syncthreads();
            }
            double t = __dadd_rn(acc, __shfl_down_sync(0xffffffffu, acc, 1));      // lanes 0,2,4,6: p0+p1, p2+p3, ...
            t = __dadd_rn(t, __shfl_down_sync(0xffffffffu, t, 2));                 // lanes 0,4: (p0+p1)+(p2+p3), ...
            t = __dadd_rn(t, __shfl_down_sync(0xffffffffu, t, 4));                 // lane 0: the canonical score
            if (valid && l8 == 0) {
                const unsigned long long e = buf[i];
                rows[i] = (uint32_t)e;
                exact[i] = t;
                my_err = fmaxf(my_err, fabsf((float)((double)__uint_as_float((uint32_t)(e >> 32)) - t)));
            }
        }
    }
    if (p.err_max) atomicMax(reinterpret_cast<int*>(&s_err), __float_as_int(my_err));       // non-negative floats order as ints
    __syncthreads();

    // ---- 3. rank by (exact desc, row asc) and emit: all-pairs rank count
    const int kk = min(p.k, n);
    for (int i = tid; i < n; i += F2_THREADS) {
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
    for (int i = kk + tid; i < p.k; i += F2_THREADS) {
        p.out_scores[(size_t)q * p.k + i] = -CUDART_INF;
        p.out_ids[(size_t)q * p.k + i] = -1;
    }
    __syncthreads();
    if (tid == 0) {
        const double q2 = s_q2[1] + s_q2[2] + s_q2[3];
        const double eps = 2.0 * (double)p.dim * 1.1920928955078125e-07 * sqrt(q2) * p.row_norm_bound;
        const float m = s_m;
        // proven complete iff nothing was dropped (m = -inf) or the k-th exact score clears m + eps
        const bool proven = p.approx || (m == -CUDART_INF_F) || (n >= p.k && s_ek_sh > (double)m + eps);
        const int flag = proven ? 0 : 1;
        if (p.out_flags) p.out_flags[q] = flag | (p.q_list ? 2 : 0);      // bit 1: the query went through the tier-0 re-scan
        if (flag) p.flag_list[atomicAdd(p.n_flagged, 1)] = q;
        if (p.err_max) p.err_max[q] = s_err;
    }
}

int launch_finish2(const FinishParams& fp, int dtype, cudaStream_t st) {
    const size_t smem = finish2_smem_bytes(fp.dim, fp.kprime);
    if (dtype == B200RAG_F16) {
        B200_CUDA_CHECK(cudaFuncSetAttribute(dense_finish2_kernel<B200RAG_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dense_finish2_kernel<B200RAG_F16><<<fp.n_q, F2_THREADS, smem, st>>>(fp); count_launch();
    } else {
        B200_CUDA_CHECK(cudaFuncSetAttribute(dense_finish2_kernel<B200RAG_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dense_finish2_kernel<B200RAG_BF16><<<fp.n_q, F2_THREADS, smem, st>>>(fp); count_launch();
    }
    B200_CUDA_CHECK(cudaGetLastError());
    return B200RAG_OK;
}

// ----------------------------------------------------------------------------------------------- split finish
// finish3 = dense_finish2_kernel in split mode (selection only) + the two kernels below.  The monolithic kernel re-scores a
// query's k' candidate rows 16 at a time inside the query's CTA: k'/16 dependent gather round trips per query, and with 256
// queries x k' = 640 only 256 CTAs to hide them.  Here the re-score is its own grid over (16-row batch, query): every
// candidate row of the whole batch is in flight at once (B = 1024, k' = 128: 8192 CTAs; B = 256, k' = 640: 10240).
template <int DTYPE>
__global__ void __launch_bounds__(F2_THREADS) finish3_rescore_kernel(const FinishParams p) {
    extern __shared__ __align__(16) char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slot_q = blockIdx.y, i0 = blockIdx.x * F2_ROWS;
    const int n = p.fin_n[slot_q];
    if (i0 >= n) return;
    double* qd = reinterpret_cast<double*>(smem);                                         // [dim] the query in fp64
    char* stage = reinterpret_cast<char*>(qd + p.dim);
    const unsigned long long* sel = p.fin_sel + (size_t)slot_q * p.kprime;
    for (int d = tid; d < p.dim; d += F2_THREADS) qd[d] = bits_to_double<DTYPE>(p.queries[(size_t)slot_q * p.dim + d]);
    const int rows_here = min(F2_ROWS, n - i0);
    const int l8 = tid & 7, grp = tid >> 3;
    const int seg = finish2_seg(p.dim);
    const int stride = seg * 2 + 16;
    const double* qj = qd + l8;
    const int i = i0 + grp;
    const bool valid = i < n;
    const uint16_t* x = reinterpret_cast<const uint16_t*>(stage + (size_t)grp * stride) + l8;
    double acc = 0.0;
    for (int d0 = 0; d0 < p.dim; d0 += seg) {                     // (a lane's chain runs on across the segments, in increasing d)
        const int dl = min(seg, p.dim - d0);
        for (int r = warp; r < rows_here; r += F2_THREADS / 32) {
            const uint32_t row = (uint32_t)sel[i0 + r];
            const uint4* src = reinterpret_cast<const uint4*>(p.corpus + (size_t)row * p.dim + d0);
            const uint32_t dst = smem_u32(stage + (size_t)r * stride);
            for (int c = lane; c < dl / 8; c += 32)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + c * 16), "l"(src + c) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                          // (also publishes qd on the first round)
        if (valid) {
#pragma unroll 4
            for (int d = 0; d < dl; d += 8) acc = fma(qj[d0 + d], bits_to_double<DTYPE>(x[d]), acc);
        }
        __syncthreads();
    }
    double t = __dadd_rn(acc, __shfl_down_sync(0xffffffffu, acc, 1));
    t = __dadd_rn(t, __shfl_down_sync(0xffffffffu, t, 2));
    t = __dadd_rn(t, __shfl_down_sync(0xffffffffu, t, 4));
    if (valid && l8 == 0) {
        p.fin_exact[(size_t)slot_q * p.kprime + i] = t;
        if (p.err_max) {
            const float err = fabsf((float)((double)__uint_as_float((uint32_t)(sel[i] >> 32)) - t));
            atomicMax(reinterpret_cast<int*>(p.err_max + slot_q), __float_as_int(err));       // non-negative floats order as ints
        }
    }
}

constexpr int F3_RANK_THREADS = 256;

template <int DTYPE>
__global__ void __launch_bounds__(F3_RANK_THREADS) finish3_rank_kernel(const FinishParams p) {
    extern __shared__ __align__(16) char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = blockIdx.x;
    double* exact = reinterpret_cast<double*>(smem);                                      // [kprime]
    uint32_t* rows = reinterpret_cast<uint32_t*>(exact + p.kprime);                       // [kprime]
    __shared__ double s_q2[F3_RANK_THREADS / 32];
    __shared__ double s_ek_sh;
    const int n = p.fin_n[q];
    for (int i = tid; i < n; i += F3_RANK_THREADS) {
        exact[i] = p.fin_exact[(size_t)q * p.kprime + i];
        rows[i] = (uint32_t)p.fin_sel[(size_t)q * p.kprime + i];
    }
    double q2 = 0.0;
    for (int d = tid; d < p.dim; d += F3_RANK_THREADS) {
        const double v = bits_to_double<DTYPE>(p.queries[(size_t)q * p.dim + d]);
        q2 = fma(v, v, q2);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) q2 += __shfl_xor_sync(0xffffffffu, q2, off);
    if (lane == 0) s_q2[warp] = q2;
    if (tid == 0) s_ek_sh = -CUDART_INF;
    __syncthreads();
    const int kk = min(p.k, n);
    for (int i = tid; i < n; i += F3_RANK_THREADS) {
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
    for (int i = kk + tid; i < p.k; i += F3_RANK_THREADS) {
        p.out_scores[(size_t)q * p.k + i] = -CUDART_INF;
        p.out_ids[(size_t)q * p.k + i] = -1;
    }
    __syncthreads();
    if (tid == 0) {
        double q2s = 0.0;                                  // (the proof's epsilon only needs |q| to a few ulps: any summation order)
        for (int w = 0; w < F3_RANK_THREADS / 32; ++w) q2s += s_q2[w];
        const double eps = 2.0 * (double)p.dim * 1.1920928955078125e-07 * sqrt(q2s) * p.row_norm_bound;
        const float m = p.fin_m[q];
        const bool proven = (m == -CUDART_INF_F) || (n >= p.k && s_ek_sh > (double)m + eps);
        const int flag = proven ? 0 : 1;
        if (p.out_flags) p.out_flags[q] = flag;
        if (flag) p.flag_list[atomicAdd(p.n_flagged, 1)] = q;
    }
}

size_t finish3_workspace_bytes(int n_q, int kprime) {
    return align_up((size_t)n_q * kprime * 8, 256) * 2 + align_up((size_t)n_q * 4, 256) * 2;
}

int launch_finish3(FinishParams fp, int dtype, void* fin_ws, cudaStream_t st) {
    char* w = static_cast<char*>(fin_ws);
    const size_t plane = align_up((size_t)fp.n_q * fp.kprime * 8, 256), small = align_up((size_t)fp.n_q * 4, 256);
    fp.fin_sel = reinterpret_cast<unsigned long long*>(w);
    fp.fin_exact = reinterpret_cast<double*>(w + plane);
    fp.fin_n = reinterpret_cast<int*>(w + 2 * plane);
    fp.fin_m = reinterpret_cast<float*>(w + 2 * plane + small);
    int rc = launch_finish2(fp, dtype, st);                       // split mode: selection only
    if (rc) return rc;
    const size_t smem_b = (size_t)fp.dim * 8 + (size_t)F2_ROWS * ((size_t)finish2_seg(fp.dim) * 2 + 16) + 16;
    const size_t smem_c = (size_t)fp.kprime * 12 + 16;
    const dim3 grid_b((fp.kprime + F2_ROWS - 1) / F2_ROWS, fp.n_q);
    if (dtype == B200RAG_F16) {
        B200_CUDA_CHECK(cudaFuncSetAttribute(finish3_rescore_kernel<B200RAG_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
        finish3_rescore_kernel<B200RAG_F16><<<grid_b, F2_THREADS, smem_b, st>>>(fp); count_launch();
        finish3_rank_kernel<B200RAG_F16><<<fp.n_q, F3_RANK_THREADS, smem_c, st>>>(fp); count_launch();
    } else {
        B200_CUDA_CHECK(cudaFuncSetAttribute(finish3_rescore_kernel<B200RAG_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
        finish3_rescore_kernel<B200RAG_BF16><<<grid_b, F2_THREADS, smem_b, st>>>(fp); count_launch();
        finish3_rank_kernel<B200RAG_BF16><<<fp.n_q, F3_RANK_THREADS, smem_c, st>>>(fp); count_launch();
    }
    B200_CUDA_CHECK(cudaGetLastError());
    return B200RAG_OK;
}

// ----------------------------------------------------------------------------------------------- sample threshold
// After the strided SAMPLE pass every (chunk, query) slot holds TC_SAMPLE_R group maxima; the rank-th greatest of a query's
// n_chunks * TC_SAMPLE_R values becomes its starting threshold (dense_tc.cu explains the sizing).  One WARP per query: the
// values are staged in shared memory as order-preserving keys and the rank-th greatest is found with a 32-step bitwise
// search (count of keys >= candidate, one warp reduction per bit).  Round 1 ran a 256-thread CTA with the block-level
// streaming top-k for this.
constexpr int ST2_WARPS = 4;

__global__ void __launch_bounds__(ST2_WARPS * 32)
sample_threshold2_kernel(const unsigned long long* __restrict__ cand, int cap, int nqb, int n_chunks, int rank, int n_q_pad,
                         unsigned int* __restrict__ gthr, const int* __restrict__ gate) {
    extern __shared__ __align__(16) char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = blockIdx.x * ST2_WARPS + warp;
    if (q >= n_q_pad) return;
    if (gate && q >= 2 * TC_BM * ((__ldg(gate) + 2 * TC_BM - 1) / (2 * TC_BM))) return;     // tier 0: pairs of query blocks that exist
    const int qb = q / TC_BM, ql = q % TC_BM;
    const int total = n_chunks * TC_SAMPLE_R;
    uint32_t* keys = reinterpret_cast<uint32_t*>(smem) + (size_t)warp * total;
    for (int e = lane; e < total; e += 32) {
        const int chunk = e / TC_SAMPLE_R, j = e % TC_SAMPLE_R;
        const unsigned long long v = __ldg(cand + (((size_t)(chunk * nqb + qb)) * TC_BM + ql) * cap + j);
        keys[e] = mono32(__uint_as_float((uint32_t)(v >> 32)));
    }
    __syncwarp();
    uint32_t T = 0;
    for (int bit = 31; bit >= 0; --bit) {
        const uint32_t c0 = T | (1u << bit);
        int c = 0;
        for (int e = lane; e < total; e += 32) c += keys[e] >= c0;
        c = __reduce_add_sync(0xffffffffu, c);
        if (c >= rank) T = c0;
    }
    if (lane == 0) gthr[q] = total >= rank ? T : 0u;
}

int launch_sample_threshold2(const unsigned long long* cand, int cap, int nqb, int n_chunks, int rank, unsigned int* gthr,
                             cudaStream_t st, const int* gate) {
    const int n_q_pad = nqb * TC_BM;
    const size_t smem = (size_t)ST2_WARPS * n_chunks * TC_SAMPLE_R * 4;
    B200_CUDA_CHECK(cudaFuncSetAttribute(sample_threshold2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sample_threshold2_kernel<<<(n_q_pad + ST2_WARPS - 1) / ST2_WARPS, ST2_WARPS * 32, smem, st>>>(cand, cap, nqb, n_chunks, rank, n_q_pad, gthr, gate);
    count_launch();
    B200_CUDA_CHECK(cudaGetLastError());
    return B200RAG_OK;
}

}  // namespace b200rag
