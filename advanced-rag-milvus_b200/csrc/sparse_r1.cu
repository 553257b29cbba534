// sparse_r1.cu -- the round-1 sparse scan, kept as a measured baseline and as the A/B partner of the round-2 kernels
// (option "sparse_flags" bit 4 = 16 selects it; tools/sparse_ab.py).  One 512-thread CTA per query walks 16384-document blocks:
// accumulators in shared memory, a block barrier between terms, candidates found by reading the accumulators back
// (profiles/r1_sparse_ncu.md).  See sparse_mask.cu / sparse_bm25.cu for what replaced it and why.
#include "sparse.cuh"

namespace b200rag {

int launch_merge(const double* cand_scores, const int64_t* cand_ids, int n_launch, const int32_t* q_list, int n_cand, int k,
                 double* out_scores_f64, float* out_scores_f32, int64_t* out_ids, int32_t* out_counts, cudaStream_t st);

// Debug: per-CTA cycle counters by phase (b200rag_debug_r1_stats_unused).  Thread 0 keeps them in shared memory (no registers).
constexpr int R1_NSTAT = 12;
constexpr int R1_STAT_CTAS = 1024;
#define R1_MARK(i)                                                                     \
    do {                                                                               \
        if (stats && tid == 0) {                                                       \
            const long long now_ = clock64();                                          \
            s_stat[i] += (unsigned long long)(now_ - s_last);                          \
            s_last = now_;                                                             \
        }                                                                              \
    } while (0)

constexpr int R1_THREADS = 512;     // with 16384-document blocks two CTAs fit per SM and overlap each other's latencies
constexpr int R1_TG = 8;       // query terms fetched together (one register pair per term and thread)

__device__ __forceinline__ void r1_cp_async8(void* smem_dst, const void* gsrc, unsigned src_bytes) {
    // 8-byte asynchronous global -> shared copy; src_bytes = 0 writes zeros (nothing is read)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc),
                 "r"(src_bytes)
                 : "memory");
}

__global__ void __launch_bounds__(R1_THREADS, 2)
sparse_r1_kernel(const int64_t* __restrict__ blk_term_ptr, const uint16_t* __restrict__ post_doc,
                    const float* __restrict__ post_w, int64_t n_docs, int n_terms, int block_docs, int n_blocks, int n_slices,
                    const int64_t* __restrict__ q_ptr, const int32_t* __restrict__ q_terms, const float* __restrict__ q_vals,
                    int k, int cap, int64_t id_offset, double* __restrict__ part_scores, int64_t* __restrict__ part_ids,
                    const uint32_t* __restrict__ doc_mask, int flags, unsigned long long* __restrict__ stats) {
    extern __shared__ __align__(16) char smem[];
    __shared__ unsigned long long s_stat[R1_NSTAT];
    __shared__ long long s_last;
    // posting ranges [s_beg, s_end) of the query's terms inside a block: slots 0..2 = the first R1_TG terms of block
    // (blk % 3), filled two blocks ahead by cp.async; slot 3 = four later terms at a time (queries with more than R1_TG
    // terms), filled synchronously
    __shared__ __align__(16) long long s_beg[4][R1_TG], s_end[4][R1_TG];
    __shared__ float s_qv[2][R1_TG];
    __shared__ int s_total[2];                                                       // candidates of block (blk & 1)
    const int tid = threadIdx.x;
    if (stats && tid == 0) {
        for (int i = 0; i < R1_NSTAT; ++i) s_stat[i] = 0;
        s_last = clock64();
    }
    const bool allow_dense = flags & 1, allow_bulk = flags & 2;
    const int q = blockIdx.x;
    const int slice = blockIdx.y;
    const int n_words = block_docs / 32;                                            // <= 2048
    float* acc = reinterpret_cast<float*>(smem);                                    // [block_docs]
    uint32_t* touched = reinterpret_cast<uint32_t*>(smem + (size_t)block_docs * 4);  // [n_words]
    char* p = smem + (size_t)block_docs * 4 + (size_t)n_words * 4;
    p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(p) + 15) & ~uintptr_t(15));
    BlockTopK<R1_THREADS, uint32_t> tk;
    tk.attach(p, cap, k, R1_THREADS, /*start_digit=*/BlockTopK<R1_THREADS, uint32_t>::NLO + 3);
    tk.init();
    for (int i = tid; i < block_docs; i += R1_THREADS) acc[i] = 0.0f;
    for (int i = tid; i < n_words; i += R1_THREADS) touched[i] = 0u;
    if (tid == 0) { s_total[0] = 0; s_total[1] = 0; }

    const int64_t qs = q_ptr[q];
    const int nq = (int)(q_ptr[q + 1] - qs);
    const int b0 = (int)((int64_t)slice * n_blocks / n_slices), b1 = (int)((int64_t)(slice + 1) * n_blocks / n_slices);
    // The chain block -> term pointers -> postings is two dependent trips to HBM, and nothing else in a block is long
    // enough to hide one.  Both are taken ahead of time:
    //   * lanes 0..R1_TG-1 copy the ranges of block blk+2 straight into shared memory (cp.async: no registers, nothing
    //     waits on it) at the top of block blk;
    //   * the postings travel in two register sets of R1_TG/2 terms: set A (terms 0..3) of block blk+1 is requested in
    //     the middle of block blk and is in flight during the rest of the block and its whole collect; set B (terms 4..7)
    //     is requested at the top of its block and has the four term steps of set A to arrive.
    int my_t = -1;
    if (tid < R1_TG) {
        float qv = 0.f;
        if (tid < nq) {
            const int t = q_terms[qs + tid];
            if (t >= 0 && t < n_terms) { my_t = t; qv = q_vals[qs + tid]; }
        }
        s_qv[0][tid] = qv;
    }
    auto stage_ranges = [&](int blk, int slot) {          // lanes 0..R1_TG-1
        const int64_t* src = blk_term_ptr + (size_t)blk * (n_terms + 1) + (my_t >= 0 ? my_t : 0);
        const unsigned sz = my_t >= 0 ? 8u : 0u;
        r1_cp_async8(&s_beg[slot][tid], src, sz);
        r1_cp_async8(&s_end[slot][tid], src + 1, sz);
    };
    constexpr int R1_H = R1_TG / 2;
    auto load_half = [&](int (&d)[R1_H], float (&w)[R1_H], int slot, int first) {
#pragma unroll
        for (int j = 0; j < R1_H; ++j) {
            const long long i = s_beg[slot][first + j] + tid;
            d[j] = -1;
            w[j] = 0.f;
            if (i < s_end[slot][first + j]) { d[j] = post_doc[i]; w[j] = post_w[i]; }
        }
    };
    const bool walk = nq > 0 && b0 < b1;
    if (tid < R1_TG && walk) {
        stage_ranges(b0, 0);
        if (b0 + 1 < b1) stage_ranges(b0 + 1, 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    int dA[R1_H];
    float wA[R1_H];
#pragma unroll
    for (int j = 0; j < R1_H; ++j) { dA[j] = -1; wA[j] = 0.f; }
    if (walk) load_half(dA, wA, 0, 0);
    R1_MARK(0);                                           // init
    bool acc_busy = false;      // CTA-uniform: the previous block's survivors may still be read out of (and zeroed in) `acc`
    int sl_cur = 0, sl_nxt = 1, sl_nx2 = 2;
    for (int blk = b0; blk < b1; ++blk) {
        const int64_t doc0 = (int64_t)blk * block_docs;
        const int64_t* tp = blk_term_ptr + (size_t)blk * (n_terms + 1);
        const int cur = blk & 1;
        float thr_f = tk.threshold_hi32_as_float();      // stable here: the previous collect ended on a barrier
        const bool dense = allow_dense && thr_f > 0.0f;  // CTA-uniform
        // one term: first R1_THREADS postings from registers, the rest of a long list straight from memory (two in flight)
        auto apply = [&](int d, float w, float qv, long long beg, long long e) {
            if (d >= 0) {
                acc[d] = fmaf(qv, w, acc[d]);
                if (!dense) atomicOr(&touched[d >> 5], 1u << (d & 31));
            }
            for (long long i = beg + tid + R1_THREADS; i < e; i += 2 * R1_THREADS) {
                const long long i1 = i + R1_THREADS;
                const int d0 = post_doc[i];
                const float w0 = post_w[i];
                int d1 = -1;
                float w1 = 0.f;
                if (i1 < e) { d1 = post_doc[i1]; w1 = post_w[i1]; }
                acc[d0] = fmaf(qv, w0, acc[d0]);
                if (!dense) atomicOr(&touched[d0 >> 5], 1u << (d0 & 31));
                if (d1 >= 0) {
                    acc[d1] = fmaf(qv, w1, acc[d1]);
                    if (!dense) atomicOr(&touched[d1 >> 5], 1u << (d1 & 31));
                }
            }
        };
        // ---- accumulate, in ascending term order with a barrier after every term ---------------------------------------
        if (nq > 0) {
            if (tid < R1_TG) {
                // slot sl_nx2 held block blk-1: last read before that block's final term barrier
                if (blk + 2 < b1) stage_ranges(blk + 2, sl_nx2);
                asm volatile("cp.async.commit_group;" ::: "memory");          // (one group per block, possibly empty)
            }
            int dB[R1_H];
            float wB[R1_H];
            load_half(dB, wB, sl_cur, R1_H);
            if (acc_busy) {                               // (waits while the postings are in flight)
                __syncthreads();
                acc_busy = false;
            }
#pragma unroll
            for (int j = 0; j < R1_H; ++j) {
                apply(dA[j], wA[j], s_qv[0][j], s_beg[sl_cur][j], s_end[sl_cur][j]);
                // the ranges of block blk+1 were requested a whole block ago: everything but the newest group has landed
                if (j == R1_H - 1 && tid < R1_TG) asm volatile("cp.async.wait_group 1;" ::: "memory");
                __syncthreads();
                if (j == 0) R1_MARK(2);                   // first term applied
            }
            if (blk + 1 < b1) load_half(dA, wA, sl_nxt, 0);
            if (nq > R1_H) {
#pragma unroll
                for (int j = 0; j < R1_H; ++j) {
                    apply(dB[j], wB[j], s_qv[0][R1_H + j], s_beg[sl_cur][R1_H + j], s_end[sl_cur][R1_H + j]);
                    __syncthreads();
                }
            }
            for (int g0 = R1_TG; g0 < nq; g0 += R1_H) {   // queries with more than R1_TG terms: four more at a time, unpipelined
                if (tid < R1_H) {
                    long long s = 0, e = 0;
                    float qv = 0.f;
                    if (g0 + tid < nq) {
                        const int t = q_terms[qs + g0 + tid];
                        if (t >= 0 && t < n_terms) { s = tp[t]; e = tp[t + 1]; qv = q_vals[qs + g0 + tid]; }
                    }
                    s_beg[3][tid] = s; s_end[3][tid] = e; s_qv[1][tid] = qv;
                }
                __syncthreads();
                load_half(dB, wB, 3, 0);
#pragma unroll
                for (int j = 0; j < R1_H; ++j) {
                    apply(dB[j], wB[j], s_qv[1][j], s_beg[3][j], s_end[3][j]);
                    __syncthreads();
                }
            }
            R1_MARK(3);                                   // remaining terms applied
            const int t_ = sl_cur; sl_cur = sl_nxt; sl_nxt = sl_nx2; sl_nx2 = t_;
        }
        // ---- collect -------------------------------------------------------------------------------------------------
        // `m` = this thread's candidate positions.  Losers are dropped with ONE float compare against the running k-th best
        // score; the exact (score, id) comparison happens only for the few candidates at or above it.
        unsigned long long m = 0;
        if (dense) {
            // bit 4*j + c  <->  document 4 * (j * R1_THREADS + tid) + c
            float4* acc4 = reinterpret_cast<float4*>(acc);
            const int nv = block_docs >> 2;
            int sh = 0;
            for (int v0 = tid; v0 < nv; v0 += 4 * R1_THREADS) {              // four LDS.128 in flight
                float4 x[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int v = v0 + u * R1_THREADS;
                    x[u] = v < nv ? acc4[v] : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u, sh += 4) {
                    const float4 y = x[u];
                    if ((__float_as_uint(y.x) | __float_as_uint(y.y) | __float_as_uint(y.z) | __float_as_uint(y.w)) == 0u) continue;
                    const unsigned b = (!(y.x < thr_f) ? 1u : 0u) | (!(y.y < thr_f) ? 2u : 0u) | (!(y.z < thr_f) ? 4u : 0u) |
                                       (!(y.w < thr_f) ? 8u : 0u);
                    float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (b) {
                        m |= (unsigned long long)b << sh;
                        if (b & 1u) z.x = y.x;
                        if (b & 2u) z.y = y.y;
                        if (b & 4u) z.z = y.z;
                        if (b & 8u) z.w = y.w;
                    }
                    acc4[v0 + u * R1_THREADS] = z;
                }
            }
        } else {
            // bits 0..31 <-> word tid of the bitmap, bits 32..63 <-> word tid + R1_THREADS
            if (tid < n_words) { m = touched[tid]; touched[tid] = 0u; }
            if (tid + R1_THREADS < n_words) { m |= (unsigned long long)touched[tid + R1_THREADS] << 32; touched[tid + R1_THREADS] = 0u; }
            if (doc_mask && m) {
                // metadata filter: documents that are not allowed are dropped here (their accumulators still have to go back
                // to zero).  block_docs is a multiple of 32, so a bitmap word of the block is a word of the mask.
                unsigned long long allowed = 0;
                const int64_t w0 = (doc0 >> 5) + tid, w1 = w0 + R1_THREADS, n_mask_words = (n_docs + 31) >> 5;
                if (tid < n_words && w0 < n_mask_words) allowed = __ldg(doc_mask + w0);
                if (tid + R1_THREADS < n_words && w1 < n_mask_words) allowed |= (unsigned long long)__ldg(doc_mask + w1) << 32;
                unsigned long long drop = m & ~allowed;
                while (drop) {
                    const int bpos = __ffsll((long long)drop) - 1;
                    drop &= drop - 1;
                    acc[bpos < 32 ? tid * 32 + bpos : (tid + R1_THREADS) * 32 + (bpos - 32)] = 0.0f;
                }
                m &= allowed;
            }
        }
        // next candidate of this thread at or above the threshold (and allowed): true + its key, or false with m == 0
        auto next_candidate = [&](const BlockTopK<R1_THREADS, uint32_t>::View& tv, uint64_t& h, uint32_t& l) -> bool {
            while (m) {
                const int bpos = __ffsll((long long)m) - 1;
                m &= m - 1;
                const int d = dense ? ((((bpos >> 2) * R1_THREADS + tid) << 2) | (bpos & 3))
                                    : (bpos < 32 ? tid * 32 + bpos : (tid + R1_THREADS) * 32 + (bpos - 32));
                const float sc = acc[d];
                acc[d] = 0.0f;
                if (sc < thr_f) continue;
                if (dense && doc_mask) {            // (the bitmap path filtered its words above)
                    const int64_t g = doc0 + d;
                    if (!((__ldg(doc_mask + (g >> 5)) >> (g & 31)) & 1u)) continue;
                }
                h = (uint64_t)mono32(sc);
                l = ~(uint32_t)(doc0 + d);
                if (tk.passes(tv, h, l)) return true;
            }
            return false;
        };
        {
            int c = __popcll(m);
            c = __reduce_add_sync(0xffffffffu, c);
            if ((tid & 31) == 0 && c) atomicAdd(&s_total[cur], c);
        }
        int held = tk.count();                            // nobody appends between the last settle and the next barrier
        __syncthreads();
        const int total = s_total[cur];
        if (tid == 0) s_total[cur ^ 1] = 0;               // the previous block's counter: its readers are barriers behind
        R1_MARK(4);                                       // accumulators scanned
        if (total == 0) continue;
        if (held + total > cap && held > k) {
            // no room for this block's survivors: keep the k best now (raises the threshold, so fewer of them survive)
            tk.compact();
            held = tk.count();
            thr_f = tk.threshold_hi32_as_float();
            if (stats && tid == 0) s_stat[9] += 1;
            R1_MARK(1);                                   // compaction
        }
        if (allow_bulk && held + total <= cap) {
            // everything fits: every thread appends all its survivors at once.  The barrier that must separate this from the
            // next block's accumulation is taken there, under the postings' latency; the count and the threshold are next
            // read behind the term barriers.
            const auto tv = tk.view();
            if (!dense || doc_mask) {
                // the dense scan kept exactly the scores >= thr_f; the bitmap walk and the document filter still have to drop theirs
                unsigned long long keep = 0;
                while (m) {
                    const int bpos = __ffsll((long long)m) - 1;
                    m &= m - 1;
                    const int d = dense ? ((((bpos >> 2) * R1_THREADS + tid) << 2) | (bpos & 3))
                                        : (bpos < 32 ? tid * 32 + bpos : (tid + R1_THREADS) * 32 + (bpos - 32));
                    bool ok = !(acc[d] < thr_f);
                    if (ok && dense && doc_mask) {      // (the bitmap path filtered its words above)
                        const int64_t g = doc0 + d;
                        ok = (__ldg(doc_mask + (g >> 5)) >> (g & 31)) & 1u;
                    }
                    if (ok) keep |= 1ull << bpos;
                    else acc[d] = 0.0f;
                }
                m = keep;
            }
            // one slot reservation per warp, then every lane moves its survivors out of `acc` on its own
            const int c = __popcll(m);
            int incl = c;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, off);
                if ((tid & 31) >= off) incl += v;
            }
            const int n_warp = __shfl_sync(0xffffffffu, incl, 31);
            if (n_warp) {
                int slot = tk.reserve_warp(n_warp) + incl - c;
                while (m) {
                    const int bpos = __ffsll((long long)m) - 1;
                    m &= m - 1;
                    const int d = dense ? ((((bpos >> 2) * R1_THREADS + tid) << 2) | (bpos & 3))
                                        : (bpos < 32 ? tid * 32 + bpos : (tid + R1_THREADS) * 32 + (bpos - 32));
                    const float sc = acc[d];
                    acc[d] = 0.0f;
                    tk.put(tv, slot++, (uint64_t)mono32(sc), ~(uint32_t)(doc0 + d));
                }
            }
            if (stats && tid == 0) { s_stat[7] += 1; s_stat[10] += total; }
            acc_busy = true;
            R1_MARK(5);                                   // bulk append
        } else {
            while (__syncthreads_or(m != 0ull)) {
                if (stats && tid == 0) s_stat[7] += 1;                  // candidate rounds
                uint64_t h = 0;
                uint32_t l = 0;
                const auto tv = tk.view();
                tk.append(tv, next_candidate(tv, h, l), h, l);
                tk.settle();
                thr_f = tk.threshold_hi32_as_float();
            }
            R1_MARK(8);                                   // candidate rounds (offer + settle, one candidate per thread)
        }
    }
    __syncthreads();
    tk.finalize();
    const int n = tk.count();
    const uint64_t* oh = tk.out_hi();
    const uint32_t* ol = tk.out_lo();
    double* ps = part_scores + ((size_t)q * n_slices + slice) * k;
    int64_t* pi = part_ids + ((size_t)q * n_slices + slice) * k;
    for (int i = tid; i < k; i += R1_THREADS) {
        if (i < n) {
            ps[i] = (double)unmono32((uint32_t)oh[i]);
            pi[i] = id_offset + (int64_t)(~ol[i]);
        } else {
            ps[i] = -CUDART_INF;
            pi[i] = -1;
        }
    }
    R1_MARK(6);                                           // finalize + output
    if (stats && tid == 0) {
        const int cta = blockIdx.y * gridDim.x + blockIdx.x;
        if (cta < R1_STAT_CTAS)
            for (int i = 0; i < R1_NSTAT; ++i) stats[(size_t)cta * R1_NSTAT + i] = s_stat[i];
    }
}

static size_t sparse_r1_smem(int block_docs, int k, int* cap_out) {
    int cap = BlockTopK<R1_THREADS, uint32_t>::capacity_for(k, R1_THREADS);
    if (cap_out) *cap_out = cap;
    return (size_t)block_docs * 4 + (size_t)((block_docs + 31) / 32) * 4 + 16 +
           BlockTopK<R1_THREADS, uint32_t>::smem_bytes(cap) + 64;
}


int launch_sparse_r1(const SparseParams& p, int n_queries, cudaStream_t st) {
    int cap = 0;
    const size_t smem = sparse_r1_smem(p.block_docs, p.k, &cap);
    if (smem > 225 * 1024) {
        set_error("sparse_topk(r1): block_docs=%d with k=%d needs %zu bytes of shared memory", p.block_docs, p.k, smem);
        return B200RAG_E_UNSUPPORTED;
    }
    B200_CUDA_CHECK(cudaFuncSetAttribute(sparse_r1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)n_queries, (unsigned)p.n_slices);
    sparse_r1_kernel<<<grid, R1_THREADS, smem, st>>>(p.blk_term_ptr, p.post_doc, p.post_w, p.n_docs, p.n_terms, p.block_docs, p.n_blocks,
                                                      p.n_slices, p.q_ptr, p.q_terms, p.q_vals, p.k, cap, p.id_offset, p.part_scores,
                                                      p.part_ids, p.doc_mask, 3, nullptr); count_launch();
    B200_CUDA_CHECK(cudaGetLastError());
    return B200RAG_OK;
}

}  // namespace b200rag
