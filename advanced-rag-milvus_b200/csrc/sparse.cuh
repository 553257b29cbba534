// sparse.cuh -- parameters of the experimental term-mask sparse kernel (sparse_mask.cu: queries of up to 15 terms, A/B only)
// and the interface between it and the product kernel's host code (sparse_bm25.cu).
#pragma once
#include "common.cuh"
#include "select.cuh"

namespace b200rag {

constexpr int SPM_MAX_TERMS = 15;   // terms per query the mask kernel handles (bit 15 of a document's mask = "scored")

struct SparseParams {
    const int64_t* blk_term_ptr;
    const uint16_t* post_doc;
    const float* post_w;
    int64_t n_docs;
    int n_terms, block_docs, n_blocks, n_slices;
    const int64_t* q_ptr;
    const int32_t* q_terms;
    const float* q_vals;
    int k, cap;
    int64_t id_offset;
    double* part_scores;        // [n_queries][n_slices][k]   (n_slices > 1)
    int64_t* part_ids;
    float* out_scores;          // [n_queries][k]             (n_slices == 1: written directly)
    int64_t* out_ids;
    int32_t* out_counts;
    unsigned int* gthr;         // [n_queries] mono32 keys of the best k-th score any slice has established (n_slices > 1)
    const uint32_t* doc_mask;
    int flags;                  // A/B: bit 0 bitmap collect only, bit 1 no staging, bit 2 no warp-private path, bit 3 general kernel only
    int mask_max_terms;         // queries with at most this many terms go to sparse_mask_kernel, the others to sparse_query_kernel
    unsigned long long* stats;
};


size_t sparse_mask_smem(int block_docs, int k, int* cap_out);
int launch_sparse_mask(const SparseParams& p, int n_queries, cudaStream_t st);

}  // namespace b200rag
