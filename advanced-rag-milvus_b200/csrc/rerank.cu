// rerank.cu -- the tail of the retrieval path on the device: learned re-rank and result-set diversity  (SURVEY 8f-4).
//
//   b200rag_rerank_learned    HybridRetriever.rerank with the LearnedRanker (reference src/advanced_rag/retrieval.py:518-563,
//                             ranker.py:109-125):  s = base_weight*score + method_bonus*len(retrieval_methods) + recency_weight*recency
//                             in fp64 in the reference's operation order, then a STABLE descending sort (Python's list.sort with
//                             reverse=True keeps equal keys in their original order) and the first rerank_top_k.
//   b200rag_pairwise_jaccard  RAGEvaluator._calculate_pairwise_similarity (reference evaluation.py:327-344): mean token-set Jaccard
//                             over the pairs i < j of a result list whose two token sets are non-empty; the mean is numpy's
//                             (np.mean = pairwise summation, restated below) so the value is bit-identical.  Same token sets as
//                             the MMR kernel (sorted unique token ids per document).
// One CTA per query; the lists are short (<= MAX_TOP_K = 100 results), so everything is all-pairs work in shared memory.
#include "common.cuh"

namespace b200rag {

constexpr int RR_THREADS = 128;
constexpr int RR_MAX = 1024;       // results per query

__global__ void __launch_bounds__(RR_THREADS)
rerank_learned_kernel(const double* __restrict__ scores, const int32_t* __restrict__ method_mask, const double* __restrict__ recency,
                      const int32_t* __restrict__ n_in, int t_max, double base_weight, double method_bonus, double recency_weight,
                      int k_out, int32_t* __restrict__ out_pos, double* __restrict__ out_scores, int32_t* __restrict__ out_n) {
    extern __shared__ __align__(16) char smem[];
    double* s = reinterpret_cast<double*>(smem);               // [t_max] learned scores
    const int q = blockIdx.x, tid = threadIdx.x;
    const int n = min(n_in[q], t_max);
    for (int i = tid; i < n; i += RR_THREADS) {
        const size_t at = (size_t)q * t_max + i;
        const double cnt = (double)__popc((unsigned)method_mask[at]);
        // Python: base_weight * base_score + method_bonus * method_count + recency_weight * recency   (left to right, no FMA)
        double v = __dadd_rn(__dmul_rn(base_weight, scores[at]), __dmul_rn(method_bonus, cnt));
        v = __dadd_rn(v, __dmul_rn(recency_weight, recency ? recency[at] : 0.0));
        s[i] = v;
    }
    __syncthreads();
    const int kk = min(k_out, n);
    for (int i = tid; i < n; i += RR_THREADS) {
        const double v = s[i];
        int rk = 0;
        for (int j = 0; j < n; ++j) rk += (s[j] > v) || (s[j] == v && j < i);      // stable: earlier wins a tie
        if (rk < kk) {
            out_pos[(size_t)q * k_out + rk] = i;
            out_scores[(size_t)q * k_out + rk] = v;
        }
    }
    for (int i = kk + tid; i < k_out; i += RR_THREADS) {
        out_pos[(size_t)q * k_out + i] = -1;
        out_scores[(size_t)q * k_out + i] = -CUDART_INF;
    }
    if (tid == 0) out_n[q] = kk;
}

// numpy's pairwise summation of a float64 array (numpy/core/src/umath/loops_utils.h.src: DOUBLE_pairwise_sum, block size 128):
//   n < 8: sequential;  n <= 128: eight interleaved accumulators, combined as a tree, then the tail;  else split at
//   n/2 rounded down to a multiple of 8 and add the two halves.  The recursion is unrolled into an explicit post-order walk (a
// recursive __device__ function needs more stack than the default per-thread limit at 4950 pairs).
__device__ double np_pairwise_leaf(const double* a, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res = __dadd_rn(res, a[i]);
        return res;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], a[i + j]);
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, a[i]);
    return res;
}

__device__ double np_pairwise_sum(const double* a, int n) {
    struct Frame { int off, len, stage; double left; };
    Frame st[24];
    int sp = 0;
    double ret = 0.0;
    st[sp++] = Frame{0, n, 0, 0.0};
    while (sp > 0) {
        Frame& f = st[sp - 1];
        if (f.stage == 0) {
            if (f.len <= 128) {
                ret = np_pairwise_leaf(a + f.off, f.len);
                --sp;
                continue;
            }
            int n2 = f.len / 2;
            n2 -= n2 % 8;
            f.stage = 1;
            st[sp++] = Frame{f.off, n2, 0, 0.0};
        } else if (f.stage == 1) {
            int n2 = f.len / 2;
            n2 -= n2 % 8;
            f.left = ret;
            f.stage = 2;
            st[sp++] = Frame{f.off + n2, f.len - n2, 0, 0.0};
        } else {
            ret = __dadd_rn(f.left, ret);
            --sp;
        }
    }
    return ret;
}

__global__ void __launch_bounds__(RR_THREADS)
pairwise_jaccard_kernel(const int32_t* __restrict__ docs, const int32_t* __restrict__ n_in, int n_max,
                        const int64_t* __restrict__ doc_tok_ptr, const int32_t* __restrict__ doc_tok_ids,
                        double* __restrict__ pair_ws, double* __restrict__ out_mean, int32_t* __restrict__ out_pairs) {
    const int q = blockIdx.x, tid = threadIdx.x;
    const int n = min(n_in[q], n_max);
    const int n_pairs = n * (n - 1) / 2;
    double* sims = pair_ws + (size_t)q * (n_max * (n_max - 1) / 2 + 1);     // this query's pair similarities, in (i, j) order
    const int32_t* dq = docs + (size_t)q * n_max;
    // pair p <-> (i, j), i < j, enumerated as the reference's nested loops do; pairs with an empty set are skipped there, so
    // the kept similarities are compacted afterwards in the same order
    for (int pidx = tid; pidx < n_pairs; pidx += RR_THREADS) {
        int i = 0, rem = pidx;
        while (rem >= n - 1 - i) { rem -= n - 1 - i; ++i; }
        const int j = i + 1 + rem;
        const int64_t a0 = doc_tok_ptr[dq[i]], a1 = doc_tok_ptr[dq[i] + 1], b0 = doc_tok_ptr[dq[j]], b1 = doc_tok_ptr[dq[j] + 1];
        double sim = -1.0;                                                  // -1 = pair skipped (an empty token set)
        if (a1 > a0 && b1 > b0) {
            int inter = 0;
            int64_t x = a0, y = b0;                                          // both lists are sorted unique token ids
            while (x < a1 && y < b1) {
                const int tx = __ldg(doc_tok_ids + x), ty = __ldg(doc_tok_ids + y);
                inter += tx == ty;
                x += tx <= ty;
                y += ty <= tx;
            }
            const int uni = (int)(a1 - a0) + (int)(b1 - b0) - inter;
            sim = __ddiv_rn((double)inter, (double)uni);
        }
        sims[pidx] = sim;
    }
    __syncthreads();
    if (tid == 0) {
        int m = 0;
        for (int pidx = 0; pidx < n_pairs; ++pidx) {
            const double v = sims[pidx];
            if (v >= 0.0) sims[m++] = v;
        }
        out_pairs[q] = m;
        out_mean[q] = m ? __ddiv_rn(np_pairwise_sum(sims, m), (double)m) : 0.0;
    }
}

}  // namespace b200rag

using namespace b200rag;

extern "C" {

int b200rag_rerank_learned(const double* scores, const int32_t* method_mask, const double* recency, const int32_t* n_in,
                           int32_t n_queries, int32_t t_max, double base_weight, double method_bonus, double recency_weight,
                           int32_t k_out, int32_t* out_pos, double* out_scores, int32_t* out_n, void* stream) {
    B200_REQUIRE(scores && method_mask && n_in && out_pos && out_scores && out_n, "rerank_learned: null pointer");
    B200_REQUIRE(n_queries >= 0 && t_max >= 1 && t_max <= RR_MAX && k_out >= 1, "rerank_learned: bad sizes (t_max <= %d)", RR_MAX);
    if (n_queries == 0) return B200RAG_OK;
    rerank_learned_kernel<<<n_queries, RR_THREADS, (size_t)t_max * 8, static_cast<cudaStream_t>(stream)>>>(
        scores, method_mask, recency, n_in, t_max, base_weight, method_bonus, recency_weight, k_out, out_pos, out_scores, out_n);
    count_launch();
    B200_CUDA_CHECK(cudaGetLastError());
    return B200RAG_OK;
}

size_t b200rag_pairwise_jaccard_workspace_bytes(int32_t n_queries, int32_t n_max) {
    if (n_queries <= 0 || n_max <= 0) return 256;
    return align_up((size_t)n_queries * ((size_t)n_max * (n_max - 1) / 2 + 1) * sizeof(double), 256) + 256;
}

int b200rag_pairwise_jaccard(const int32_t* docs, const int32_t* n_in, int32_t n_queries, int32_t n_max,
                             const int64_t* doc_tok_ptr, const int32_t* doc_tok_ids, double* out_mean, int32_t* out_pairs,
                             void* workspace, size_t workspace_bytes, void* stream) {
    B200_REQUIRE(docs && n_in && doc_tok_ptr && out_mean && out_pairs && workspace, "pairwise_jaccard: null pointer");
    B200_REQUIRE(n_queries >= 0 && n_max >= 1 && n_max <= RR_MAX, "pairwise_jaccard: bad sizes (n_max <= %d)", RR_MAX);
    if (n_queries == 0) return B200RAG_OK;
    if (workspace_bytes < b200rag_pairwise_jaccard_workspace_bytes(n_queries, n_max)) {
        set_error("pairwise_jaccard: workspace too small (%zu < %zu)", workspace_bytes, b200rag_pairwise_jaccard_workspace_bytes(n_queries, n_max));
        return B200RAG_E_WORKSPACE;
    }
    pairwise_jaccard_kernel<<<n_queries, RR_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        docs, n_in, n_max, doc_tok_ptr, doc_tok_ids, static_cast<double*>(workspace), out_mean, out_pairs);
    count_launch();
    B200_CUDA_CHECK(cudaGetLastError());
    return B200RAG_OK;
}

}  // extern "C"
