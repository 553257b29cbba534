// sparse_bm25.cu -- sparse inner-product (BM25-weighted) top-k over doc-range-blocked postings  (K3).
//
// One CTA (512 threads) owns one QUERY and walks the document blocks of its slice one after the other:
//   * the block's per-document accumulators live in shared memory (fp32, block_docs <= 32768 -> 128 KB) and are zeroed
//     once per CTA; collecting a block's candidates puts them back to zero;
//   * the postings of up to 8 query terms inside the block are fetched TOGETHER (one coalesced u16 doc + f32 weight per
//     thread and term; the ranges come from a shared-memory ring filled two blocks ahead with cp.async, the first four
//     terms of the NEXT block are requested in the middle of the current one) and then applied in ascending term id with
//     a barrier between terms, so every document sees  acc = fmaf(qv, w, acc)  in the canonical order (bit-identical to
//     oracle/exact_scan.c:orc_sparse_topk).  Postings of one term hit distinct documents: no atomics on the accumulators;
//   * candidates are collected in one of two ways, chosen per block from the running k-th best score `thr`:
//       thr > 0  (the steady state): every thread reads 32+ accumulators with LDS.128, keeps the few that are >= thr and
//                writes zeros back -- no bitmap, no atomics in the accumulate step (an untouched document holds exactly 0,
//                so it can never pass a positive threshold);
//       otherwise (first block(s), or < k positive candidates so far): a touched-bitmap marks the candidates and a
//                thread walks the set bits of one or two 32-document words;
//     survivors go to the block-level streaming top-k (select.cuh), which lives for the whole walk, so later blocks
//     are filtered by the threshold earlier blocks established: appended in bulk (one slot reservation per warp, no
//     barrier until the next block's postings are in flight) whenever they fit, compacted only when the buffer is full.
// Algorithmic HBM traffic = 6 bytes per posting of the query's terms.  grid = (queries, slices): with fewer queries
// than SMs the blocks are split into slices; merge_topk_kernel reduces the slices.
//
// Round 2 measured four replacements for this kernel on the same box (tools/sparse_ab.py, profiles/r2_sparse_experiments.md):
// ordered per-term phases sized by list length + an exchange collect proportional to the postings, the same with the next
// block's postings staged by cp.async, warp-private document sub-ranges (no block barrier between terms), and a kernel
// without accumulators at all (sparse_mask.cu, kept selectable).  All four returned bit-identical results and all four were
// SLOWER (0.85 - 1.4 ms against 0.65 ms at config 4): with ~135 postings per (block, term) list the work is too fine-grained
// for any of those structures to amortise their per-list and per-block bookkeeping -- they executed 320 - 400 M warp
// instructions where this kernel needs 273 M.  This kernel therefore stays the product path.
#include "sparse.cuh"

namespace b200rag {

int launch_merge(const double* cand_scores, const int64_t* cand_ids, int n_launch, const int32_t* q_list, int n_cand, int k,
                 double* out_scores_f64, float* out_scores_f32, int64_t* out_ids, int32_t* out_counts, cudaStream_t st);

// Debug: per-CTA cycle counters by phase (b200rag_debug_r1_stats_unused).  Thread 0 keeps them in shared memory (no registers).
constexpr int SP_NSTAT = 12;
constexpr int SP_STAT_CTAS = 1024;
#define SP_MARK(i)                                                                     \
    do {                                                                               \
        if (stats && tid == 0) {                                                       \
            const long long now_ = clock64();                                          \
            s_stat[i] += (unsigned long long)(now_ - s_last);                          \
            s_last = now_;                                                             \
        }                                                                              \
    } while (0)

constexpr int SP_THREADS = 512;     // with 16384-document blocks two CTAs fit per SM and overlap each other's latencies
constexpr int SP_TG = 8;       // query terms fetched together (one register pair per term and thread)

__device__ __forceinline__ void sp_cp_async8(void* smem_dst, const void* gsrc, unsigned src_bytes) {
    // 8-byte asynchronous global -> shared copy; src_bytes = 0 writes zeros (nothing is read)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc),
                 "r"(src_bytes)
                 : "memory");
}

__global__ void __launch_bounds__(SP_THREADS, 2)
sparse_query_kernel(const int64_t* __restrict__ blk_term_ptr, const uint16_t* __restrict__ post_doc,
                    const float* __restrict__ post_w, int64_t n_docs, int n_terms, int block_docs, int n_blocks, int n_slices,
                    const int64_t* __restrict__ q_ptr, const int32_t* __restrict__ q_terms, const float* __restrict__ q_vals,
                    int k, int cap, int64_t id_offset, double* __restrict__ part_scores, int64_t* __restrict__ part_ids,
                    const uint32_t* __restrict__ doc_mask, int flags, unsigned long long* __restrict__ stats, int skip_upto_terms) {
    if ((int)(q_ptr[blockIdx.x + 1] - q_ptr[blockIdx.x]) <= skip_upto_terms) return;   // served by sparse_mask_kernel (A/B mode)
    extern __shared__ __align__(16) char smem[];
    __shared__ unsigned long long s_stat[SP_NSTAT];
    __shared__ long long s_last;
    // posting ranges [s_beg, s_end) of the query's terms inside a block: slots 0..2 = the first SP_TG terms of block
    // (blk % 3), filled two blocks ahead by cp.async; slot 3 = four later terms at a time (queries with more than SP_TG
    // terms), filled synchronously
    __shared__ __align__(16) long long s_beg[4][SP_TG], s_end[4][SP_TG];
    __shared__ float s_qv[2][SP_TG];
    __shared__ int s_total[2];                                                       // candidates of block (blk & 1)
    const int tid = threadIdx.x;
    if (stats && tid == 0) {
        for (int i = 0; i < SP_NSTAT; ++i) s_stat[i] = 0;
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
    BlockTopK<SP_THREADS, uint32_t> tk;
    tk.attach(p, cap, k, SP_THREADS, /*start_digit=*/BlockTopK<SP_THREADS, uint32_t>::NLO + 3);
    tk.init();
    for (int i = tid; i < block_docs; i += SP_THREADS) acc[i] = 0.0f;
    for (int i = tid; i < n_words; i += SP_THREADS) touched[i] = 0u;
    if (tid == 0) { s_total[0] = 0; s_total[1] = 0; }

    const int64_t qs = q_ptr[q];
    const int nq = (int)(q_ptr[q + 1] - qs);
    const int b0 = (int)((int64_t)slice * n_blocks / n_slices), b1 = (int)((int64_t)(slice + 1) * n_blocks / n_slices);
    // The chain block -> term pointers -> postings is two dependent trips to HBM, and nothing else in a block is long
    // enough to hide one.  Both are taken ahead of time:
    //   * lanes 0..SP_TG-1 copy the ranges of block blk+2 straight into shared memory (cp.async: no registers, nothing
    //     waits on it) at the top of block blk;
    //   * the postings travel in two register sets of SP_TG/2 terms: set A (terms 0..3) of block blk+1 is requested in
    //     the middle of block blk and is in flight during the rest of the block and its whole collect; set B (terms 4..7)
    //     is requested at the top of its block and has the four term steps of set A to arrive.
    int my_t = -1;
    if (tid < SP_TG) {
        float qv = 0.f;
        if (tid < nq) {
            const int t = q_terms[qs + tid];
            if (t >= 0 && t < n_terms) { my_t = t; qv = q_vals[qs + tid]; }
        }
        s_qv[0][tid] = qv;
    }
    auto stage_ranges = [&](int blk, int slot) {          // lanes 0..SP_TG-1
        const int64_t* src = blk_term_ptr + (size_t)blk * (n_terms + 1) + (my_t >= 0 ? my_t : 0);
        const unsigned sz = my_t >= 0 ? 8u : 0u;
        sp_cp_async8(&s_beg[slot][tid], src, sz);
        sp_cp_async8(&s_end[slot][tid], src + 1, sz);
    };
    constexpr int SP_H = SP_TG / 2;
    auto load_half = [&](int (&d)[SP_H], float (&w)[SP_H], int slot, int first) {
#pragma unroll
        for (int j = 0; j < SP_H; ++j) {
            const long long i = s_beg[slot][first + j] + tid;
            d[j] = -1;
            w[j] = 0.f;
            if (i < s_end[slot][first + j]) { d[j] = post_doc[i]; w[j] = post_w[i]; }
        }
    };
    const bool walk = nq > 0 && b0 < b1;
    if (tid < SP_TG && walk) {
        stage_ranges(b0, 0);
        if (b0 + 1 < b1) stage_ranges(b0 + 1, 1);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    int dA[SP_H];
    float wA[SP_H];
#pragma unroll
    for (int j = 0; j < SP_H; ++j) { dA[j] = -1; wA[j] = 0.f; }
    if (walk) load_half(dA, wA, 0, 0);
    SP_MARK(0);                                           // init
    bool acc_busy = false;      // CTA-uniform: the previous block's survivors may still be read out of (and zeroed in) `acc`
    int sl_cur = 0, sl_nxt = 1, sl_nx2 = 2;
    for (int blk = b0; blk < b1; ++blk) {
        const int64_t doc0 = (int64_t)blk * block_docs;
        const int64_t* tp = blk_term_ptr + (size_t)blk * (n_terms + 1);
        const int cur = blk & 1;
        float thr_f = tk.threshold_hi32_as_float();      // stable here: the previous collect ended on a barrier
        const bool dense = allow_dense && thr_f > 0.0f;  // CTA-uniform
        // one term: first SP_THREADS postings from registers, the rest of a long list straight from memory (two in flight)
        auto apply = [&](int d, float w, float qv, long long beg, long long e) {
            if (d >= 0) {
                acc[d] = fmaf(qv, w, acc[d]);
                if (!dense) atomicOr(&touched[d >> 5], 1u << (d & 31));
            }
            for (long long i = beg + tid + SP_THREADS; i < e; i += 2 * SP_THREADS) {
                const long long i1 = i + SP_THREADS;
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
            if (tid < SP_TG) {
                // slot sl_nx2 held block blk-1: last read before that block's final term barrier
                if (blk + 2 < b1) stage_ranges(blk + 2, sl_nx2);
                asm volatile("cp.async.commit_group;" ::: "memory");          // (one group per block, possibly empty)
            }
            int dB[SP_H];
            float wB[SP_H];
            load_half(dB, wB, sl_cur, SP_H);
            if (acc_busy) {                               // (waits while the postings are in flight)
                __syncthreads();
                acc_busy = false;
            }
#pragma unroll
            for (int j = 0; j < SP_H; ++j) {
                apply(dA[j], wA[j], s_qv[0][j], s_beg[sl_cur][j], s_end[sl_cur][j]);
                // the ranges of block blk+1 were requested a whole block ago: everything but the newest group has landed
                if (j == SP_H - 1 && tid < SP_TG) asm volatile("cp.async.wait_group 1;" ::: "memory");
                __syncthreads();
                if (j == 0) SP_MARK(2);                   // first term applied
            }
            if (blk + 1 < b1) load_half(dA, wA, sl_nxt, 0);
            if (nq > SP_H) {
#pragma unroll
                for (int j = 0; j < SP_H; ++j) {
                    apply(dB[j], wB[j], s_qv[0][SP_H + j], s_beg[sl_cur][SP_H + j], s_end[sl_cur][SP_H + j]);
                    __syncthreads();
                }
            }
            for (int g0 = SP_TG; g0 < nq; g0 += SP_H) {   // queries with more than SP_TG terms: four more at a time, unpipelined
                if (tid < SP_H) {
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
                for (int j = 0; j < SP_H; ++j) {
                    apply(dB[j], wB[j], s_qv[1][j], s_beg[3][j], s_end[3][j]);
                    __syncthreads();
                }
            }
            SP_MARK(3);                                   // remaining terms applied
            const int t_ = sl_cur; sl_cur = sl_nxt; sl_nxt = sl_nx2; sl_nx2 = t_;
        }
        // ---- collect -------------------------------------------------------------------------------------------------
        // `m` = this thread's candidate positions.  Losers are dropped with ONE float compare against the running k-th best
        // score; the exact (score, id) comparison happens only for the few candidates at or above it.
        unsigned long long m = 0;
        if (dense) {
            // bit 4*j + c  <->  document 4 * (j * SP_THREADS + tid) + c
            float4* acc4 = reinterpret_cast<float4*>(acc);
            const int nv = block_docs >> 2;
            int sh = 0;
            for (int v0 = tid; v0 < nv; v0 += 4 * SP_THREADS) {              // four LDS.128 in flight
                float4 x[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int v = v0 + u * SP_THREADS;
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
                    acc4[v0 + u * SP_THREADS] = z;
                }
            }
        } else {
            // bits 0..31 <-> word tid of the bitmap, bits 32..63 <-> word tid + SP_THREADS
            if (tid < n_words) { m = touched[tid]; touched[tid] = 0u; }
            if (tid + SP_THREADS < n_words) { m |= (unsigned long long)touched[tid + SP_THREADS] << 32; touched[tid + SP_THREADS] = 0u; }
            if (doc_mask && m) {
                // metadata filter: documents that are not allowed are dropped here (their accumulators still have to go back
                // to zero).  block_docs is a multiple of 32, so a bitmap word of the block is a word of the mask.
                unsigned long long allowed = 0;
                const int64_t w0 = (doc0 >> 5) + tid, w1 = w0 + SP_THREADS, n_mask_words = (n_docs + 31) >> 5;
                if (tid < n_words && w0 < n_mask_words) allowed = __ldg(doc_mask + w0);
                if (tid + SP_THREADS < n_words && w1 < n_mask_words) allowed |= (unsigned long long)__ldg(doc_mask + w1) << 32;
                unsigned long long drop = m & ~allowed;
                while (drop) {
                    const int bpos = __ffsll((long long)drop) - 1;
                    drop &= drop - 1;
                    acc[bpos < 32 ? tid * 32 + bpos : (tid + SP_THREADS) * 32 + (bpos - 32)] = 0.0f;
                }
                m &= allowed;
            }
        }
        // next candidate of this thread at or above the threshold (and allowed): true + its key, or false with m == 0
        auto next_candidate = [&](const BlockTopK<SP_THREADS, uint32_t>::View& tv, uint64_t& h, uint32_t& l) -> bool {
            while (m) {
                const int bpos = __ffsll((long long)m) - 1;
                m &= m - 1;
                const int d = dense ? ((((bpos >> 2) * SP_THREADS + tid) << 2) | (bpos & 3))
                                    : (bpos < 32 ? tid * 32 + bpos : (tid + SP_THREADS) * 32 + (bpos - 32));
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
        SP_MARK(4);                                       // accumulators scanned
        if (total == 0) continue;
        if (held + total > cap && held > k) {
            // no room for this block's survivors: keep the k best now (raises the threshold, so fewer of them survive)
            tk.compact();
            held = tk.count();
            thr_f = tk.threshold_hi32_as_float();
            if (stats && tid == 0) s_stat[9] += 1;
            SP_MARK(1);                                   // compaction
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
                    const int d = dense ? ((((bpos >> 2) * SP_THREADS + tid) << 2) | (bpos & 3))
                                        : (bpos < 32 ? tid * 32 + bpos : (tid + SP_THREADS) * 32 + (bpos - 32));
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
                    const int d = dense ? ((((bpos >> 2) * SP_THREADS + tid) << 2) | (bpos & 3))
                                        : (bpos < 32 ? tid * 32 + bpos : (tid + SP_THREADS) * 32 + (bpos - 32));
                    const float sc = acc[d];
                    acc[d] = 0.0f;
                    tk.put(tv, slot++, (uint64_t)mono32(sc), ~(uint32_t)(doc0 + d));
                }
            }
            if (stats && tid == 0) { s_stat[7] += 1; s_stat[10] += total; }
            acc_busy = true;
            SP_MARK(5);                                   // bulk append
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
            SP_MARK(8);                                   // candidate rounds (offer + settle, one candidate per thread)
        }
    }
    __syncthreads();
    tk.finalize();
    const int n = tk.count();
    const uint64_t* oh = tk.out_hi();
    const uint32_t* ol = tk.out_lo();
    double* ps = part_scores + ((size_t)q * n_slices + slice) * k;
    int64_t* pi = part_ids + ((size_t)q * n_slices + slice) * k;
    for (int i = tid; i < k; i += SP_THREADS) {
        if (i < n) {
            ps[i] = (double)unmono32((uint32_t)oh[i]);
            pi[i] = id_offset + (int64_t)(~ol[i]);
        } else {
            ps[i] = -CUDART_INF;
            pi[i] = -1;
        }
    }
    SP_MARK(6);                                           // finalize + output
    if (stats && tid == 0) {
        const int cta = blockIdx.y * gridDim.x + blockIdx.x;
        if (cta < SP_STAT_CTAS)
            for (int i = 0; i < SP_NSTAT; ++i) stats[(size_t)cta * SP_NSTAT + i] = s_stat[i];
    }
}

static size_t sparse_smem(int block_docs, int k, int* cap_out) {
    int cap = BlockTopK<SP_THREADS, uint32_t>::capacity_for(k, SP_THREADS);
    if (cap_out) *cap_out = cap;
    return (size_t)block_docs * 4 + (size_t)((block_docs + 31) / 32) * 4 + 16 +
           BlockTopK<SP_THREADS, uint32_t>::smem_bytes(cap) + 64;
}


}  // namespace b200rag

using namespace b200rag;

extern "C" {

size_t b200rag_sparse_topk_workspace_bytes(int64_t n_docs, int32_t block_docs, int32_t n_queries, int32_t k) {
    if (block_docs <= 0 || n_queries <= 0 || k <= 0) return 0;
    int64_t n_blocks = (n_docs + block_docs - 1) / block_docs;
    if (n_blocks < 1) n_blocks = 1;
    return align_up((size_t)n_queries * n_blocks * k * sizeof(double), 256) +
           align_up((size_t)n_queries * n_blocks * k * sizeof(int64_t), 256) + align_up((size_t)n_queries * 4, 256) + 512;
}

int b200rag_sparse_topk(const int64_t* blk_term_ptr, const uint16_t* post_doc, const float* post_w,
                        int64_t n_docs, int32_t n_terms, int32_t block_docs,
                        const int64_t* q_ptr, const int32_t* q_terms, const float* q_vals,
                        int32_t n_queries, int32_t k, int64_t id_offset,
                        float* out_scores, int64_t* out_ids, int32_t* out_counts,
                        void* workspace, size_t workspace_bytes, void* stream) {
    return b200rag_sparse_topk_masked(blk_term_ptr, post_doc, post_w, n_docs, n_terms, block_docs, q_ptr, q_terms, q_vals,
                                      n_queries, k, id_offset, out_scores, out_ids, out_counts, nullptr, workspace,
                                      workspace_bytes, stream);
}

int b200rag_sparse_topk_masked(const int64_t* blk_term_ptr, const uint16_t* post_doc, const float* post_w,
                               int64_t n_docs, int32_t n_terms, int32_t block_docs,
                               const int64_t* q_ptr, const int32_t* q_terms, const float* q_vals,
                               int32_t n_queries, int32_t k, int64_t id_offset,
                               float* out_scores, int64_t* out_ids, int32_t* out_counts, const uint32_t* doc_mask,
                               void* workspace, size_t workspace_bytes, void* stream) {
    B200_REQUIRE(blk_term_ptr && q_ptr && out_scores && out_ids && out_counts && workspace, "sparse_topk: null pointer");
    B200_REQUIRE(n_docs >= 0 && n_terms > 0 && n_queries >= 0 && k > 0, "sparse_topk: bad sizes");
    B200_REQUIRE(block_docs > 0 && block_docs <= 64 * SP_THREADS && block_docs % 32 == 0,
                 "sparse_topk: block_docs must be a multiple of 32 in (0, %d], got %d", 64 * SP_THREADS, block_docs);
    if (n_queries == 0) return B200RAG_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int cap = 0;
    const size_t smem = sparse_smem(block_docs, k, &cap);
    if (smem > 225 * 1024) {
        set_error("sparse_topk: block_docs=%d with k=%d needs %zu bytes of shared memory (max 230400); use a smaller block",
                  block_docs, k, smem);
        return B200RAG_E_UNSUPPORTED;
    }
    int64_t n_blocks = (n_docs + block_docs - 1) / block_docs;
    if (n_blocks < 1) n_blocks = 1;
    B200_REQUIRE(n_queries <= 2147483647 && n_blocks <= 2147483647, "sparse_topk: too many blocks");
    // one CTA per query; with fewer queries than ~2 per SM the blocks are cut into slices to fill the machine
    int sm_count = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    int64_t n_slices = (2 * (int64_t)sm_count) / n_queries;
    if (n_slices < 1) n_slices = 1;
    // A/B switches (b200rag_set_option; tools/sparse_ab.py): "sparse_slices" = slices per query; "sparse_flags" bit 0 dense collect
    // mode, bit 1 bulk append (default 3), bit 3 (8) = the term-mask kernel (sparse_mask.cu) for queries of up to 15 terms
    if (option(OPT_SPARSE_SLICES, 0) > 0) n_slices = option(OPT_SPARSE_SLICES, 0);
    const int ab = option(OPT_SPARSE_FLAGS, 3);
    const int flags = ab & 3;
    if (n_slices > n_blocks) n_slices = n_blocks;
    B200_REQUIRE(n_slices <= 65535, "sparse_topk: too many slices");
    Workspace ws(workspace, workspace_bytes);
    double* part_scores = ws.take<double>((size_t)n_queries * n_slices * k);
    int64_t* part_ids = ws.take<int64_t>((size_t)n_queries * n_slices * k);
    unsigned int* gthr = ws.take<unsigned int>((size_t)n_queries);
    if (!ws.ok()) {
        set_error("sparse_topk: workspace too small (%zu < %zu)", workspace_bytes, ws.off);
        return B200RAG_E_WORKSPACE;
    }
    unsigned long long* stats = stats_buffer(STATS_SPARSE, (size_t)SP_STAT_CTAS * SP_NSTAT);
    if (stats) B200_CUDA_CHECK(cudaMemsetAsync(stats, 0, (size_t)SP_STAT_CTAS * SP_NSTAT * 8, st));
    int skip_upto = -1;
    if ((ab & 8) && sparse_mask_smem(block_docs, k, nullptr) <= 225 * 1024 && ((uintptr_t)post_doc & 15) == 0 &&
        ((uintptr_t)post_w & 15) == 0 && block_docs % 8 == 0) {
        // A/B: the accumulator-free kernel for the queries it can serve; this file's kernel takes the longer ones.  How long the
        // queries are is only known on the device, so both are launched and the CTAs that do not serve their query leave at once.
        SparseParams p;
        p.blk_term_ptr = blk_term_ptr; p.post_doc = post_doc; p.post_w = post_w; p.n_docs = n_docs; p.n_terms = n_terms;
        p.block_docs = block_docs; p.n_blocks = (int)n_blocks; p.n_slices = (int)n_slices; p.q_ptr = q_ptr; p.q_terms = q_terms;
        p.q_vals = q_vals; p.k = k; p.cap = 0; p.id_offset = id_offset; p.part_scores = part_scores; p.part_ids = part_ids;
        p.out_scores = nullptr; p.out_ids = nullptr; p.out_counts = nullptr; p.doc_mask = doc_mask; p.flags = ab;
        p.mask_max_terms = SPM_MAX_TERMS; p.stats = stats;
        p.gthr = n_slices > 1 ? gthr : nullptr;
        if (n_slices > 1) B200_CUDA_CHECK(cudaMemsetAsync(gthr, 0, (size_t)n_queries * 4, st));
        int rc = launch_sparse_mask(p, n_queries, st);
        if (rc) return rc;
        skip_upto = SPM_MAX_TERMS;
    }
    B200_CUDA_CHECK(cudaFuncSetAttribute(sparse_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)n_queries, (unsigned)n_slices);
    sparse_query_kernel<<<grid, SP_THREADS, smem, st>>>(blk_term_ptr, post_doc, post_w, n_docs, n_terms, block_docs,
                                                       (int)n_blocks, (int)n_slices, q_ptr, q_terms, q_vals, k, cap, id_offset,
                                                       part_scores, part_ids, doc_mask, flags, skip_upto < 0 ? stats : nullptr,
                                                       skip_upto); count_launch();
    B200_CUDA_CHECK(cudaGetLastError());
    return launch_merge(part_scores, part_ids, n_queries, nullptr, (int)(n_slices * k), k, nullptr, out_scores, out_ids,
                        out_counts, st);
}

}  // extern "C"
