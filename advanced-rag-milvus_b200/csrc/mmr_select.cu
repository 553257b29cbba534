// mmr_select.cu -- greedy Maximal Marginal Relevance on token-set Jaccard for a batch of queries  (K6).
//
// Restates HybridRetriever._mmr_diversify (reference src/advanced_rag/retrieval.py:493-516):
//   pick 1: argmax rel;  pick j>1: argmax  lambda*rel - (1-lambda)*max_{s in selected} J(c, s)
//   J(a,b) = |a & b| / (|a | b| or 1) on token sets; strict '>' => the earliest candidate (fused order) wins ties.
// fp64 with explicit round-to-nearest ops in Python's evaluation order; the running max over the selected set is
// kept incrementally (max is order independent), turning the reference's O(k^2 n) set rebuilds into O(k n)
// intersections.
//
// One CTA per query, up to 1024 threads.  Token sets are rows of a CSR (sorted unique token ids per document).  Per pick:
//   1. argmax over the alive candidates (thread-strided, warp shuffle + one cross-warp step; "earliest wins" on ties)
//   2. the picked document's tokens are raised in a shared-memory bitset over the vocabulary
//   3. every WARP intersects whole candidates against the bitset: the 32 lanes read 32 consecutive tokens of the
//      candidate's list (one coalesced 128-byte load), probe the bitset, and the counts are summed with one warp
//      reduction.  (Round 1 had one THREAD per candidate walking its list token by token: 26.7 ms for 256 queries x 1000
//      candidates x 100 picks; this form: see profiles/r1_hybrid_c4.md.)
//   4. the bitset is cleared again (only the picked document's words).
// Work = picks x candidates x tokens per candidate bitset probes; the shared-memory bank conflicts of the random probes
// (~3.5 wavefronts per 32 probes) are the bound.
#include "common.cuh"

namespace b200rag {

constexpr int MMR_MAX_THREADS = 1024;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ bool mmr_better(double ob, int oi, double b, int bi) {
    return oi != 0x7fffffff && (bi == 0x7fffffff || ob > b || (ob == b && oi < bi));
}

__global__ void __launch_bounds__(MMR_MAX_THREADS)
mmr_select_kernel(const int32_t* __restrict__ cand_doc, const double* __restrict__ cand_rel, const int32_t* __restrict__ cand_n,
                  int n_max, const int64_t* __restrict__ doc_tok_ptr, const int32_t* __restrict__ doc_tok_ids, int vocab_words,
                  const double* __restrict__ lambda, const int32_t* __restrict__ k_sel, int k_max,
                  int32_t* __restrict__ out_pick, int32_t* __restrict__ out_n) {
    extern __shared__ __align__(16) char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthreads = blockDim.x, nwarps = nthreads >> 5;
    const int q = blockIdx.x;
    double* rel = reinterpret_cast<double*>(smem);                 // [n_max]
    double* max_sim = rel + n_max;                                 // [n_max]
    int64_t* tok_begin = reinterpret_cast<int64_t*>(max_sim + n_max);   // [n_max] start of the candidate's token list
    int* tok_len = reinterpret_cast<int*>(tok_begin + n_max);      // [n_max]
    int* alive = tok_len + n_max;                                  // [n_max]
    uint32_t* bits = reinterpret_cast<uint32_t*>(alive + n_max);   // [vocab_words]
    __shared__ double s_best[MMR_MAX_THREADS / 32];
    __shared__ int s_best_idx[MMR_MAX_THREADS / 32];
    __shared__ int s_pick;
    __shared__ int s_done;

    const int n = min(cand_n[q], n_max);
    const int k = min(min(k_sel[q], k_max), n);
    const double lam = lambda[q];
    const double one_minus = __dsub_rn(1.0, lam);
    const int32_t* docs = cand_doc + (size_t)q * n_max;
    for (int i = tid; i < n; i += nthreads) {
        rel[i] = cand_rel[(size_t)q * n_max + i];
        max_sim[i] = 0.0;
        alive[i] = 1;
        const int64_t b0 = doc_tok_ptr[docs[i]];
        tok_begin[i] = b0;
        tok_len[i] = (int)(doc_tok_ptr[docs[i] + 1] - b0);
    }
    for (int i = tid; i < vocab_words; i += nthreads) bits[i] = 0u;
    if (tid == 0) s_done = 0;
    __syncthreads();

    for (int step = 0; step < k; ++step) {
        // ---- 1. argmax with "earliest wins" -------------------------------------------------------
        double best = -1e9;
        int best_i = 0x7fffffff;
        for (int c = tid; c < n; c += nthreads) {
            if (!alive[c]) continue;
            double s = step == 0 ? rel[c] : __dsub_rn(__dmul_rn(lam, rel[c]), __dmul_rn(one_minus, max_sim[c]));
            if (s > best) { best = s; best_i = c; }   // per-thread candidates come in increasing c
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            double ob = __shfl_down_sync(FULL, best, off);
            int oi = __shfl_down_sync(FULL, best_i, off);
            if (mmr_better(ob, oi, best, best_i)) { best = ob; best_i = oi; }
        }
        if (lane == 0) { s_best[warp] = best; s_best_idx[warp] = best_i; }
        __syncthreads();
        if (warp == 0) {
            double b = lane < nwarps ? s_best[lane] : -1e9;
            int bi = lane < nwarps ? s_best_idx[lane] : 0x7fffffff;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                double ob = __shfl_down_sync(FULL, b, off);
                int oi = __shfl_down_sync(FULL, bi, off);
                if (mmr_better(ob, oi, b, bi)) { b = ob; bi = oi; }
            }
            if (lane == 0) {
                s_pick = bi;
                if (bi != 0x7fffffff) {
                    out_pick[(size_t)q * k_max + step] = bi;
                    alive[bi] = 0;
                    s_done = step + 1;
                }
            }
        }
        __syncthreads();
        const int pick = s_pick;
        if (pick == 0x7fffffff) break;           // nothing beat -1e9 (the reference would fail here too)
        if (step + 1 == k) break;
        // ---- 2. raise the picked document's tokens -------------------------------------------------
        const int64_t ps = tok_begin[pick];
        const int len_p = tok_len[pick];
        for (int i = tid; i < len_p; i += nthreads) {
            int t = doc_tok_ids[ps + i];
            atomicOr(&bits[t >> 5], 1u << (t & 31));
        }
        __syncthreads();
        // ---- 3. warp per candidate: |tokens(c) & tokens(pick)| -------------------------------------
        for (int c = warp; c < n; c += nwarps) {
            if (!alive[c]) continue;              // warp-uniform
            const int32_t* toks = doc_tok_ids + tok_begin[c];
            const int len_c = tok_len[c];
            int inter = 0;
            for (int i = lane; i < len_c; i += 32) {
                const int t = __ldg(toks + i);
                inter += (bits[t >> 5] >> (t & 31)) & 1u;
            }
            inter = __reduce_add_sync(FULL, inter);
            if (lane == 0) {
                const int uni = len_c + len_p - inter;
                const double j = __ddiv_rn((double)inter, (double)(uni ? uni : 1));
                if (j > max_sim[c]) max_sim[c] = j;
            }
        }
        __syncthreads();
        // ---- 4. clear (the next raise happens two barriers later) ----------------------------------
        for (int i = tid; i < len_p; i += nthreads) bits[doc_tok_ids[ps + i] >> 5] = 0u;
    }
    __syncthreads();
    const int done = s_done;
    if (tid == 0) out_n[q] = done;
    for (int i = done + tid; i < k_max; i += nthreads) out_pick[(size_t)q * k_max + i] = -1;
}

}  // namespace b200rag

using namespace b200rag;

extern "C" {

size_t b200rag_mmr_select_workspace_bytes(int32_t, int32_t, int32_t) { return 256; }

int b200rag_mmr_select(const int32_t* cand_doc, const double* cand_rel, const int32_t* cand_n, int32_t n_queries,
                       int32_t n_max, const int64_t* doc_tok_ptr, const int32_t* doc_tok_ids, int32_t vocab_size,
                       const double* lambda, const int32_t* k_sel, int32_t k_max,
                       int32_t* out_pick, int32_t* out_n,
                       void* workspace, size_t workspace_bytes, void* stream) {
    (void)workspace; (void)workspace_bytes;
    B200_REQUIRE(cand_doc && cand_rel && cand_n && doc_tok_ptr && lambda && k_sel && out_pick && out_n, "mmr_select: null pointer");
    B200_REQUIRE(n_queries >= 0 && n_max >= 1 && vocab_size >= 1 && k_max >= 1, "mmr_select: bad sizes");
    if (n_queries == 0) return B200RAG_OK;
    int vocab_words = (vocab_size + 31) / 32;
    size_t smem = (size_t)n_max * (8 + 8 + 8 + 4 + 4) + (size_t)vocab_words * 4 + 64;
    if (smem > 225 * 1024) {
        set_error("mmr_select: n_max=%d vocab=%d needs %zu bytes of shared memory", n_max, vocab_size, smem);
        return B200RAG_E_UNSUPPORTED;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    B200_CUDA_CHECK(cudaFuncSetAttribute(mmr_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // one warp per candidate in the intersection phase: as many warps as there are candidates, up to 32
    const int threads = n_max >= 32 ? MMR_MAX_THREADS : (n_max >= 8 ? 256 : 128);
    mmr_select_kernel<<<n_queries, threads, smem, st>>>(cand_doc, cand_rel, cand_n, n_max, doc_tok_ptr, doc_tok_ids,
                                                           vocab_words, lambda, k_sel, k_max, out_pick, out_n); count_launch();
    B200_CUDA_CHECK(cudaGetLastError());
    return B200RAG_OK;
}

}  // extern "C"
