// mmr_select.cu -- greedy Maximal Marginal Relevance on token-set Jaccard for a batch of queries  (K6).
//
// Restates HybridRetriever._mmr_diversify (reference src/advanced_rag/retrieval.py:493-516):
//   pick 1: argmax rel;  pick j>1: argmax  lambda*rel - (1-lambda)*max_{s in selected} J(c, s)
//   J(a,b) = |a & b| / (|a | b| or 1) on token sets; strict '>' => the earliest candidate (fused order) wins ties.
// fp64 with explicit round-to-nearest ops in Python's evaluation order; the running max over the selected set is
// kept incrementally (max is order independent), turning the reference's O(k^2 n) set rebuilds into O(k n)
// intersections.
//
// One CTA per query.  Token sets are rows of a CSR (sorted unique token ids per document).  Per pick: the picked
// document's tokens are raised in a shared-memory bitset over the vocabulary, every thread intersects its
// candidates against the bitset, the bitset is cleared again.
#include "common.cuh"

namespace b200rag {

constexpr int MMR_THREADS = 256;

__global__ void __launch_bounds__(MMR_THREADS)
mmr_select_kernel(const int32_t* __restrict__ cand_doc, const double* __restrict__ cand_rel, const int32_t* __restrict__ cand_n,
                  int n_max, const int64_t* __restrict__ doc_tok_ptr, const int32_t* __restrict__ doc_tok_ids, int vocab_words,
                  const double* __restrict__ lambda, const int32_t* __restrict__ k_sel, int k_max,
                  int32_t* __restrict__ out_pick, int32_t* __restrict__ out_n) {
    extern __shared__ __align__(16) char smem[];
    const int tid = threadIdx.x;
    const int q = blockIdx.x;
    double* rel = reinterpret_cast<double*>(smem);                 // [n_max]
    double* max_sim = rel + n_max;                                 // [n_max]
    int* alive = reinterpret_cast<int*>(max_sim + n_max);          // [n_max]
    uint32_t* bits = reinterpret_cast<uint32_t*>(alive + n_max);   // [vocab_words]
    __shared__ double s_best[MMR_THREADS / 32];
    __shared__ int s_best_idx[MMR_THREADS / 32];
    __shared__ int s_pick;
    __shared__ int s_done;

    const int n = min(cand_n[q], n_max);
    const int k = min(min(k_sel[q], k_max), n);
    const double lam = lambda[q];
    const double one_minus = __dsub_rn(1.0, lam);
    const int32_t* docs = cand_doc + (size_t)q * n_max;
    for (int i = tid; i < n; i += MMR_THREADS) {
        rel[i] = cand_rel[(size_t)q * n_max + i];
        max_sim[i] = 0.0;
        alive[i] = 1;
    }
    for (int i = tid; i < vocab_words; i += MMR_THREADS) bits[i] = 0u;
    if (tid == 0) s_done = 0;
    __syncthreads();

    for (int step = 0; step < k; ++step) {
        // ---- argmax with "earliest wins" --------------------------------------------------------
        double best = -1e9;
        int best_i = 0x7fffffff;
        for (int c = tid; c < n; c += MMR_THREADS) {
            if (!alive[c]) continue;
            double s = step == 0 ? rel[c] : __dsub_rn(__dmul_rn(lam, rel[c]), __dmul_rn(one_minus, max_sim[c]));
            if (s > best) { best = s; best_i = c; }   // per-thread candidates come in increasing c
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            double ob = __shfl_down_sync(0xffffffffu, best, off);
            int oi = __shfl_down_sync(0xffffffffu, best_i, off);
            if (oi != 0x7fffffff && (best_i == 0x7fffffff || ob > best || (ob == best && oi < best_i))) { best = ob; best_i = oi; }
        }
        if ((tid & 31) == 0) { s_best[tid >> 5] = best; s_best_idx[tid >> 5] = best_i; }
        __syncthreads();
        if (tid == 0) {
            double b = s_best[0];
            int bi = s_best_idx[0];
            for (int w = 1; w < MMR_THREADS / 32; ++w) {
                double ob = s_best[w];
                int oi = s_best_idx[w];
                if (oi != 0x7fffffff && (bi == 0x7fffffff || ob > b || (ob == b && oi < bi))) { b = ob; bi = oi; }
            }
            s_pick = bi;
            if (bi != 0x7fffffff) {
                out_pick[(size_t)q * k_max + step] = bi;
                alive[bi] = 0;
                s_done = step + 1;
            }
        }
        __syncthreads();
        const int pick = s_pick;
        if (pick == 0x7fffffff) break;           // nothing beat -1e9 (the reference would fail here too)
        if (step + 1 == k) break;
        // ---- raise the picked document's tokens, intersect, clear --------------------------------
        const int64_t ps = doc_tok_ptr[docs[pick]], pe = doc_tok_ptr[docs[pick] + 1];
        const int len_p = (int)(pe - ps);
        for (int64_t i = ps + tid; i < pe; i += MMR_THREADS) {
            int t = doc_tok_ids[i];
            atomicOr(&bits[t >> 5], 1u << (t & 31));
        }
        __syncthreads();
        for (int c = tid; c < n; c += MMR_THREADS) {
            if (!alive[c]) continue;
            const int64_t cs = doc_tok_ptr[docs[c]], ce = doc_tok_ptr[docs[c] + 1];
            int inter = 0;
            for (int64_t i = cs; i < ce; ++i) {
                int t = __ldg(doc_tok_ids + i);
                inter += (bits[t >> 5] >> (t & 31)) & 1u;
            }
            int uni = (int)(ce - cs) + len_p - inter;
            double j = __ddiv_rn((double)inter, (double)(uni ? uni : 1));
            if (j > max_sim[c]) max_sim[c] = j;
        }
        __syncthreads();
        for (int64_t i = ps + tid; i < pe; i += MMR_THREADS) {
            int t = doc_tok_ids[i];
            bits[t >> 5] = 0u;
        }
        __syncthreads();
    }
    __syncthreads();
    const int done = s_done;
    if (tid == 0) out_n[q] = done;
    for (int i = done + tid; i < k_max; i += MMR_THREADS) out_pick[(size_t)q * k_max + i] = -1;
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
    size_t smem = (size_t)n_max * (8 + 8 + 4) + (size_t)vocab_words * 4 + 64;
    if (smem > 225 * 1024) {
        set_error("mmr_select: n_max=%d vocab=%d needs %zu bytes of shared memory", n_max, vocab_size, smem);
        return B200RAG_E_UNSUPPORTED;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    B200_CUDA_CHECK(cudaFuncSetAttribute(mmr_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mmr_select_kernel<<<n_queries, MMR_THREADS, smem, st>>>(cand_doc, cand_rel, cand_n, n_max, doc_tok_ptr, doc_tok_ids,
                                                           vocab_words, lambda, k_sel, k_max, out_pick, out_n); count_launch();
    B200_CUDA_CHECK(cudaGetLastError());
    return B200RAG_OK;
}

}  // extern "C"
