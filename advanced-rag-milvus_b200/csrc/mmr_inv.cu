// mmr_inv.cu -- greedy MMR on token-set Jaccard, third kernel generation: per-query INVERTED candidate lists  (K6).
//
// Same arithmetic and tie rule as mmr_select.cu (reference src/advanced_rag/retrieval.py:493-516).  What changes is how
// |tokens(c) & tokens(pick)| is found for every candidate c after a pick.  The first two generations probe, for every alive
// candidate, each of its ~90 tokens against a bitset of the picked document: n x len probes per pick (10^5 at config 4),
// although two random documents share only ~20 tokens.  Here the candidate x token incidence of the query is transposed ONCE:
//
//   build   1. every token of every candidate is marked in a vocabulary bitmap (shared memory); a prefix popcount turns a
//              token id into a COMPACT id 0 .. D-1 (D = distinct tokens among the query's candidates, < 65536);
//           2. the candidates' lists are rewritten as compact ids (fwd, u16, global workspace) and counted per compact id;
//           3. an exclusive scan turns the counts into list starts and a scatter fills  inv[start[id] ..] = the candidates
//              that hold token id  (counting sort, candidate order inside a list is irrelevant).
//   pick    the picked document's ~90 compact ids name ~90 inverted lists; their postings -- one per (token, candidate)
//           incidence, i.e. exactly sum_c |tokens(c) & tokens(pick)| of them, ~25 per thread instead of ~90 probes -- are walked
//           as one flat index space and counted into per-candidate shared-memory counters.
//   The running max similarity is kept as an integer fraction: inter/union > num/den  <=>  inter*den > num*union, exact in
//   64-bit integers and equivalent to comparing the correctly rounded fp64 quotients (distinct fractions with denominators
//   < 2^17 differ by more than 2^-34); the fp64 division -- ~50 instructions -- runs only when a candidate's maximum changes.
//
// One 1024-thread CTA per query, thread = candidate; ~30 KB of shared memory at a 100K-token vocabulary, so two CTAs share an
// SM and 256 queries run in one wave.  Queries the scheme cannot hold (more than 65535 distinct tokens, a document of more than
// 1024 tokens, more incidences than the workspace has room for) are marked and served by the bitset kernels of mmr_select.cu,
// which are launched behind this one and leave at once for every other query.
#include "common.cuh"

namespace b200rag {

constexpr int MI_THREADS = 1024;
constexpr int MI_WARPS = MI_THREADS / 32;
constexpr unsigned MI_FULL = 0xffffffffu;
constexpr int MI_MAX_D = 65535;             // compact ids are u16
constexpr int MI_MAX_LEN = 1024;            // tokens per document (the picked document's lists are scanned by one thread each)
constexpr int MI_STAGE = 32768;             // incidences of one pick staged in shared memory per round (64 KB)
constexpr uint32_t MI_PAD = 0xffffu;        // filler of lists of odd length (lists are copied as aligned 4-byte words)

__device__ __forceinline__ bool mi_better(double ob, int oi, double b, int bi) {
    return oi != 0x7fffffff && (bi == 0x7fffffff || ob > b || (ob == b && oi < bi));
}

// exclusive block scan of one int per thread (1024 threads); returns the prefix, *total = sum.  Two barriers.
__device__ __forceinline__ int mi_block_scan(int v, int* s_wsum, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(MI_FULL, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int w = s_wsum[lane];
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(MI_FULL, wi, o);
            if (lane >= o) wi += u;
        }
        s_wsum[lane] = wi - w;
        if (lane == 31) s_wsum[32] = wi;
    }
    __syncthreads();
    *total = s_wsum[32];
    return incl - v + s_wsum[warp];
}

__global__ void __launch_bounds__(MI_THREADS, 2)
mmr_select_inv_kernel(const int32_t* __restrict__ cand_doc, const double* __restrict__ cand_rel, const int32_t* __restrict__ cand_n,
                      int n_max, const int64_t* __restrict__ doc_tok_ptr, const int32_t* __restrict__ doc_tok_ids, int vocab_words,
                      const double* __restrict__ lambda, const int32_t* __restrict__ k_sel, int k_max,
                      int32_t* __restrict__ out_pick, int32_t* __restrict__ out_n,
                      uint16_t* __restrict__ ws_fwd, uint16_t* __restrict__ ws_inv, uint32_t* __restrict__ ws_start, int t_cap) {
    extern __shared__ __align__(16) char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = blockIdx.x;
    uint32_t* bits = reinterpret_cast<uint32_t*>(smem);                          // [vocab_words] tokens present among the candidates
    uint16_t* wpre = reinterpret_cast<uint16_t*>(bits + vocab_words);             // [vocab_words] compact id of a word's first token
    uint32_t* cand_off = reinterpret_cast<uint32_t*>(wpre + vocab_words + (vocab_words & 1));   // [MI_THREADS] start of c's fwd list
    uint32_t* inter = cand_off + MI_THREADS;                                      // [MI_THREADS + 32] per-candidate intersection counters (+ a sink for the filler)
    uint32_t* s_b = inter + MI_THREADS + 32;                                           // [MI_THREADS] picked doc: begin of token t's list
    uint32_t* s_pre = s_b + MI_THREADS;                                           // [MI_THREADS] picked doc: flat prefix of list lengths
    uint16_t* stage = reinterpret_cast<uint16_t*>(s_pre + MI_THREADS);            // [MI_STAGE] the pick's inverted lists, back to back
    __shared__ int s_wsum[33];
    __shared__ double s_best[MI_WARPS];
    __shared__ int s_best_idx[MI_WARPS];
    __shared__ int s_pick, s_done, s_fail;

    uint16_t* fwd = ws_fwd + (size_t)q * t_cap;
    uint16_t* inv = ws_inv + (size_t)q * t_cap;
    uint32_t* start = ws_start + (size_t)q * (MI_MAX_D + 1);

    const int n = min(min(cand_n[q], n_max), MI_THREADS);
    const int k = min(min(k_sel[q], k_max), n);
    if (k <= 0) {                                             // nothing to pick (k_sel = 0: the caller skips this query)
        if (tid == 0) out_n[q] = 0;
        for (int i = tid; i < k_max; i += MI_THREADS) out_pick[(size_t)q * k_max + i] = -1;
        return;
    }
    const double lam = lambda[q];
    const double one_minus = __dsub_rn(1.0, lam);
    const int32_t* docs = cand_doc + (size_t)q * n_max;

    // ---- this thread's candidate
    const int c = tid;
    double rel = 0.0;
    long long tbeg = 0;
    int len = 0;
    bool alive = c < n;
    if (alive) {
        rel = cand_rel[(size_t)q * n_max + c];
        tbeg = doc_tok_ptr[docs[c]];
        len = (int)(doc_tok_ptr[docs[c] + 1] - tbeg);
    }
    for (int i = tid; i < vocab_words; i += MI_THREADS) bits[i] = 0u;
    inter[tid] = 0u;
    if (tid == 0) { s_done = 0; s_fail = 0; }
    int total_tok = 0;
    const int off = mi_block_scan(len, s_wsum, &total_tok);   // (its barriers also publish the zeroed arrays)
    cand_off[tid] = (uint32_t)off;
    if (len > MI_MAX_LEN) s_fail = 1;
    __syncthreads();
    if (total_tok > t_cap || s_fail) {                        // not representable here: mmr_select.cu's kernels take the query
        if (tid == 0) out_n[q] = -2;
        return;
    }
    // ---- build 1: mark every candidate token (warp per candidate, coalesced reads)
    for (int cc = warp; cc < n; cc += MI_WARPS) {
        const long long tb = doc_tok_ptr[docs[cc]];
        const int lc = (int)(doc_tok_ptr[docs[cc] + 1] - tb);
        for (int i = lane; i < lc; i += 32) {
            const int t = __ldg(doc_tok_ids + tb + i);
            atomicOr(&bits[t >> 5], 1u << (t & 31));
        }
    }
    __syncthreads();
    // compact id of the first token of every bitmap word = exclusive prefix of the words' popcounts
    int n_distinct = 0;
    {
        const int per = (vocab_words + MI_THREADS - 1) / MI_THREADS;
        const int w0 = tid * per, w1 = min(vocab_words, w0 + per);
        int sum = 0;
        for (int w = w0; w < w1; ++w) sum += __popc(bits[w]);
        int run = mi_block_scan(sum, s_wsum, &n_distinct);
        if (n_distinct <= MI_MAX_D) {
            for (int w = w0; w < w1; ++w) { wpre[w] = (uint16_t)run; run += __popc(bits[w]); }
        }
    }
    if (n_distinct > MI_MAX_D || total_tok + n_distinct > t_cap) {        // (every list may grow by one filler entry)
        if (tid == 0) out_n[q] = -2;
        return;
    }
    for (int i = tid; i <= n_distinct; i += MI_THREADS) start[i] = 0u;
    __syncthreads();
    // ---- build 2: compact forward lists + per-token counts (counts go to start[id + 1]: the scan below makes them list starts)
    for (int cc = warp; cc < n; cc += MI_WARPS) {
        const long long tb = doc_tok_ptr[docs[cc]];
        const int lc = (int)(doc_tok_ptr[docs[cc] + 1] - tb);
        uint16_t* dst = fwd + cand_off[cc];
        for (int i = lane; i < lc; i += 32) {
            const int t = __ldg(doc_tok_ids + tb + i);
            const uint32_t wbits = bits[t >> 5];
            const int id = (int)wpre[t >> 5] + __popc(wbits & ((1u << (t & 31)) - 1u));
            dst[i] = (uint16_t)id;
            atomicAdd(&start[id + 1], 1u);
        }
    }
    __syncthreads();
    // ---- build 3: exclusive scan of the counts in place.  start[id] = first slot of id's list; the scatter then advances
    //      start[id] to the END of the list, so afterwards list(id) = [id ? start[id - 1] : 0, start[id])
    {
        const int per = (n_distinct + MI_THREADS) / MI_THREADS;            // covers indices 1 .. n_distinct
        const int i0 = 1 + tid * per, i1 = min(n_distinct + 1, i0 + per);
        // lists are padded to an even length so that every list starts on a 4-byte boundary of inv (cp.async granularity)
        int sum = 0;
        for (int i = i0; i < i1; ++i) sum += ((int)start[i] + 1) & ~1;
        int tot = 0;
        int run = mi_block_scan(sum, s_wsum, &tot);
        // start[i] (i >= 1) holds count(i - 1); the exclusive prefix over those slots leaves begin(id) in start[id + 1]
        for (int i = i0; i < i1; ++i) {
            const int cnt = (int)start[i];
            start[i] = (uint32_t)run;
            if (cnt & 1) inv[run + cnt] = (uint16_t)MI_PAD;            // filler behind a list of odd length
            run += (cnt + 1) & ~1;
        }
    }
    __syncthreads();
    // the scatter takes its slot from start[id + 1] and advances it, so afterwards start[id + 1] = end(id) (before the filler):
    // list(id) occupies [even(start[id]), start[id + 1]) plus the filler; start[0] = 0 is never touched
    for (int cc = warp; cc < n; cc += MI_WARPS) {
        const int lc = (int)(doc_tok_ptr[docs[cc] + 1] - doc_tok_ptr[docs[cc]]);
        const uint16_t* src = fwd + cand_off[cc];
        for (int i = lane; i < lc; i += 32) {
            const int id = src[i];
            const uint32_t pos = atomicAdd(&start[id + 1], 1u);
            inv[pos] = (uint16_t)cc;
        }
    }
    __syncthreads();

    // ---- picks
    int max_num = 0, max_den = 1;                              // running max Jaccard of this candidate as a fraction
    double max_sim = 0.0;
    for (int step = 0; step < k; ++step) {
        // 1. argmax with "earliest wins"
        double best = -1e9;
        int best_i = 0x7fffffff;
        if (alive) {
            const double sc = step == 0 ? rel : __dsub_rn(__dmul_rn(lam, rel), __dmul_rn(one_minus, max_sim));
            if (sc > best) { best = sc; best_i = c; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(MI_FULL, best, o);
            const int oi = __shfl_xor_sync(MI_FULL, best_i, o);
            if (mi_better(ob, oi, best, best_i)) { best = ob; best_i = oi; }
        }
        if (lane == 0) { s_best[warp] = best; s_best_idx[warp] = best_i; }
        __syncthreads();
        if (warp == 0) {
            double b = s_best[lane];
            int bi = s_best_idx[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_xor_sync(MI_FULL, b, o);
                const int oi = __shfl_xor_sync(MI_FULL, bi, o);
                if (mi_better(ob, oi, b, bi)) { b = ob; bi = oi; }
            }
            if (lane == 0) {
                s_pick = bi;
                if (bi != 0x7fffffff) {
                    out_pick[(size_t)q * k_max + step] = bi;
                    s_done = step + 1;
                }
            }
        }
        __syncthreads();
        const int pick = s_pick;
        if (pick == 0x7fffffff) break;                        // nothing beat -1e9 (the reference would fail here too)
        if (step + 1 == k) break;
        if (c == pick) alive = false;
        // 2. the picked document's inverted lists: thread t < len_p owns token t of the pick
        const int len_p = (int)(doc_tok_ptr[docs[pick] + 1] - doc_tok_ptr[docs[pick]]);
        uint32_t lb = 0, ll = 0;
        if (tid < len_p) {
            // (ld.cg: the scatter advanced these slots with L2 atomics after this SM had read them -- its L1 may be stale)
            const int id = __ldcg(fwd + cand_off[pick] + tid);
            lb = (__ldcg(start + id) + 1u) & ~1u;             // begin(id): end(id - 1) rounded up over its filler; start[0] = 0
            ll = ((__ldcg(start + id + 1) - lb) + 1u) & ~1u;  // padded length (the filler counts into a sink slot)
        }
        int total = 0;
        const int pre = mi_block_scan((int)ll, s_wsum, &total);
        s_b[tid] = lb;
        s_pre[tid] = (uint32_t)pre;
        __syncthreads();
        // 3. the lists are copied into shared memory back to back -- aligned 4-byte cp.async copies, a warp per list, nothing
        //    waits on a load until all of them are in flight -- and then counted as one flat array: which token an entry
        //    belongs to no longer matters
        for (int r0 = 0; r0 < total; r0 += MI_STAGE) {        // (one round unless the pick shares > 32768 incidences)
            for (int t = warp; t < len_p; t += MI_WARPS) {
                const int p0 = (int)s_pre[t], p1 = t + 1 < len_p ? (int)s_pre[t + 1] : total;
                const int a0 = max(p0, r0), a1 = min(p1, r0 + MI_STAGE);          // this round's part of list t (even bounds)
                const uint32_t* src = reinterpret_cast<const uint32_t*>(inv + s_b[t] + (uint32_t)(a0 - p0));
                uint32_t* dst = reinterpret_cast<uint32_t*>(stage + (a0 - r0));
                for (int j = lane; j < (a1 - a0) / 2; j += 32)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + j)), "l"(src + j) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();
            const int n_here = min(MI_STAGE, total - r0);
            for (int w = tid; w < n_here; w += MI_THREADS) {
                const uint32_t cc = stage[w];
                atomicAdd(&inter[cc < MI_THREADS ? cc : MI_THREADS], 1u);        // (filler -> sink slot)
            }
            __syncthreads();
        }
        // 4. every candidate folds its intersection into its running maximum (integer fractions; divide only on change)
        {
            const int in = (int)inter[tid];
            inter[tid] = 0u;
            if (alive && in > 0) {
                const int uni = len + len_p - in;             // > 0 because in > 0
                if ((long long)in * max_den > (long long)max_num * uni) {
                    max_num = in;
                    max_den = uni;
                    max_sim = __ddiv_rn((double)in, (double)uni);
                }
            }
        }
        // (the next step's first barrier separates the counter reset from the next walk)
    }
    __syncthreads();
    const int done = s_done;
    if (tid == 0) out_n[q] = done;
    for (int i = done + tid; i < k_max; i += MI_THREADS) out_pick[(size_t)q * k_max + i] = -1;
}

size_t mmr_inv_smem_bytes(int vocab_words) {
    return (size_t)vocab_words * 4 + (size_t)(vocab_words + (vocab_words & 1)) * 2 + (size_t)(4 * MI_THREADS + 32) * 4 +
           (size_t)MI_STAGE * 2 + 64;
}

// Per query: forward + inverted lists (u16 each, t_cap entries) and the list starts (u32, 65536 + 1 entries).
size_t mmr_inv_workspace_bytes(int n_queries, int t_cap) {
    return align_up((size_t)n_queries * t_cap * 2, 256) * 2 + align_up((size_t)n_queries * (MI_MAX_D + 1) * 4, 256) + 256;
}

int launch_mmr_inv(const int32_t* cand_doc, const double* cand_rel, const int32_t* cand_n, int n_queries, int n_max,
                   const int64_t* doc_tok_ptr, const int32_t* doc_tok_ids, int vocab_words, const double* lambda, const int32_t* k_sel,
                   int k_max, int32_t* out_pick, int32_t* out_n, void* workspace, int t_cap, cudaStream_t st) {
    char* ws = static_cast<char*>(workspace);
    const size_t lists = align_up((size_t)n_queries * t_cap * 2, 256);
    uint16_t* fwd = reinterpret_cast<uint16_t*>(ws);
    uint16_t* inv = reinterpret_cast<uint16_t*>(ws + lists);
    uint32_t* start = reinterpret_cast<uint32_t*>(ws + 2 * lists);
    const size_t smem = mmr_inv_smem_bytes(vocab_words);
    B200_CUDA_CHECK(cudaFuncSetAttribute(mmr_select_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mmr_select_inv_kernel<<<n_queries, MI_THREADS, smem, st>>>(cand_doc, cand_rel, cand_n, n_max, doc_tok_ptr, doc_tok_ids, vocab_words,
                                                                 lambda, k_sel, k_max, out_pick, out_n, fwd, inv, start, t_cap);
    count_launch();
    B200_CUDA_CHECK(cudaGetLastError());
    return B200RAG_OK;
}

}  // namespace b200rag
