// sparse_bm25.cu -- sparse inner-product (BM25-weighted) top-k over doc-range-blocked postings  (K3).
//
// One CTA owns one (query, document block) pair.  The block's per-document accumulators live in shared memory
// (fp32, block_docs <= 65536 -> <= 256 KB is too much, so block_docs is capped by the shared memory left after the
// selection buffers; the index builder uses 32768 by default = 128 KB).  Query terms are processed in ascending
// term id with a barrier between terms, so every document's accumulator sees  acc = fmaf(qv, w, acc)  in the
// canonical order (bit-identical to oracle/exact_scan.c:orc_sparse_topk).  Within one term the postings touch
// distinct documents, so no atomics are needed on the accumulators; postings are read with coalesced loads
// (u16 doc + f32 weight, 6 bytes per posting = the algorithmic HBM traffic).  A touched-bitmap marks candidate
// documents; the final pass streams the touched documents into the block-level top-k (select.cuh) and writes a
// sorted partial list per (query, block); merge_topk_kernel reduces the blocks.
#include "common.cuh"
#include "select.cuh"

namespace b200rag {

int launch_merge(const double* cand_scores, const int64_t* cand_ids, int n_launch, const int32_t* q_list, int n_cand, int k,
                 double* out_scores_f64, float* out_scores_f32, int64_t* out_ids, int32_t* out_counts, cudaStream_t st);

constexpr int SP_THREADS = 256;
constexpr int SP_ITEMS = 4;   // documents examined per thread between two settle() calls

__global__ void __launch_bounds__(SP_THREADS)
sparse_block_kernel(const int64_t* __restrict__ blk_term_ptr, const uint16_t* __restrict__ post_doc,
                    const float* __restrict__ post_w, int64_t n_docs, int n_terms, int block_docs, int n_blocks,
                    const int64_t* __restrict__ q_ptr, const int32_t* __restrict__ q_terms, const float* __restrict__ q_vals,
                    int k, int cap, int64_t id_offset, double* __restrict__ part_scores, int64_t* __restrict__ part_ids) {
    extern __shared__ __align__(16) char smem[];
    const int tid = threadIdx.x;
    const int blk = blockIdx.x;
    const int q = blockIdx.y;
    const int64_t doc0 = (int64_t)blk * block_docs;
    const int docs_here = (int)min((int64_t)block_docs, n_docs - doc0);

    float* acc = reinterpret_cast<float*>(smem);                                   // [block_docs]
    uint32_t* touched = reinterpret_cast<uint32_t*>(smem + (size_t)block_docs * 4);  // [block_docs/32]
    const int n_words = (block_docs + 31) / 32;
    char* p = smem + (size_t)block_docs * 4 + (size_t)n_words * 4;
    p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(p) + 15) & ~uintptr_t(15));
    BlockTopK<SP_THREADS, uint32_t> tk;
    tk.attach(p, cap, k, SP_THREADS * SP_ITEMS, /*start_digit=*/BlockTopK<SP_THREADS, uint32_t>::NLO + 3);
    tk.init();
    for (int i = tid; i < block_docs; i += SP_THREADS) acc[i] = 0.0f;
    for (int i = tid; i < n_words; i += SP_THREADS) touched[i] = 0u;
    __syncthreads();

    const int64_t* tp = blk_term_ptr + (size_t)blk * (n_terms + 1);
    const int64_t qs = q_ptr[q], qe = q_ptr[q + 1];
    for (int64_t j = qs; j < qe; ++j) {
        const int t = q_terms[j];
        if (t < 0 || t >= n_terms) continue;          // uniform across the block
        const float qv = q_vals[j];
        const int64_t s = tp[t], e = tp[t + 1];
        for (int64_t i = s + tid; i < e; i += SP_THREADS) {
            const int d = post_doc[i];
            const float w = post_w[i];
            acc[d] = fmaf(qv, w, acc[d]);
            atomicOr(&touched[d >> 5], 1u << (d & 31));
        }
        __syncthreads();
    }

    // stream touched documents into the top-k
    for (int base = 0; base < n_words * 32; base += SP_THREADS * SP_ITEMS) {
#pragma unroll
        for (int it = 0; it < SP_ITEMS; ++it) {
            const int d = base + it * SP_THREADS + tid;
            bool valid = d < docs_here && ((touched[d >> 5] >> (d & 31)) & 1u);
            tk.offer(valid, valid ? (uint64_t)mono32(acc[d]) : 0, ~(uint32_t)d);
        }
        tk.settle();
    }
    __syncthreads();
    tk.finalize();
    const int n = tk.count();
    const uint64_t* oh = tk.out_hi();
    const uint32_t* ol = tk.out_lo();
    double* ps = part_scores + ((size_t)q * n_blocks + blk) * k;
    int64_t* pi = part_ids + ((size_t)q * n_blocks + blk) * k;
    for (int i = tid; i < k; i += SP_THREADS) {
        if (i < n) {
            ps[i] = (double)unmono32((uint32_t)oh[i]);
            pi[i] = id_offset + doc0 + (int64_t)(~ol[i]);
        } else {
            ps[i] = -CUDART_INF;
            pi[i] = -1;
        }
    }
}

static size_t sparse_smem(int block_docs, int k, int* cap_out) {
    int cap = BlockTopK<SP_THREADS, uint32_t>::capacity_for(k, SP_THREADS * SP_ITEMS);
    if (cap_out) *cap_out = cap;
    return (size_t)block_docs * 4 + (size_t)((block_docs + 31) / 32) * 4 + 16 +
           BlockTopK<SP_THREADS, uint32_t>::smem_bytes(cap) + 64;
}

}  // namespace b200rag

using namespace b200rag;

extern "C" {

size_t b200rag_sparse_topk_workspace_bytes(int64_t n_docs, int32_t block_docs, int32_t n_queries, int32_t k) {
    if (block_docs <= 0) return 0;
    int64_t n_blocks = (n_docs + block_docs - 1) / block_docs;
    if (n_blocks < 1) n_blocks = 1;
    return align_up((size_t)n_queries * n_blocks * k * sizeof(double), 256) +
           align_up((size_t)n_queries * n_blocks * k * sizeof(int64_t), 256) + 512;
}

int b200rag_sparse_topk(const int64_t* blk_term_ptr, const uint16_t* post_doc, const float* post_w,
                        int64_t n_docs, int32_t n_terms, int32_t block_docs,
                        const int64_t* q_ptr, const int32_t* q_terms, const float* q_vals,
                        int32_t n_queries, int32_t k, int64_t id_offset,
                        float* out_scores, int64_t* out_ids, int32_t* out_counts,
                        void* workspace, size_t workspace_bytes, void* stream) {
    B200_REQUIRE(blk_term_ptr && q_ptr && out_scores && out_ids && out_counts && workspace, "sparse_topk: null pointer");
    B200_REQUIRE(n_docs >= 0 && n_terms > 0 && n_queries >= 0 && k > 0, "sparse_topk: bad sizes");
    B200_REQUIRE(block_docs > 0 && block_docs <= 65536 && block_docs % 32 == 0,
                 "sparse_topk: block_docs must be a multiple of 32 in (0, 65536], got %d", block_docs);
    if (n_queries == 0) return B200RAG_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int cap = 0;
    size_t smem = sparse_smem(block_docs, k, &cap);
    if (smem > 225 * 1024) {
        set_error("sparse_topk: block_docs=%d with k=%d needs %zu bytes of shared memory (max 230400); use a smaller block",
                  block_docs, k, smem);
        return B200RAG_E_UNSUPPORTED;
    }
    int64_t n_blocks = (n_docs + block_docs - 1) / block_docs;
    if (n_blocks < 1) n_blocks = 1;
    B200_REQUIRE(n_blocks <= 65535 * 32, "sparse_topk: too many blocks");
    Workspace ws(workspace, workspace_bytes);
    double* part_scores = ws.take<double>((size_t)n_queries * n_blocks * k);
    int64_t* part_ids = ws.take<int64_t>((size_t)n_queries * n_blocks * k);
    if (!ws.ok()) {
        set_error("sparse_topk: workspace too small (%zu < %zu)", workspace_bytes, ws.off);
        return B200RAG_E_WORKSPACE;
    }
    B200_CUDA_CHECK(cudaFuncSetAttribute(sparse_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)n_blocks, (unsigned)n_queries);
    sparse_block_kernel<<<grid, SP_THREADS, smem, st>>>(blk_term_ptr, post_doc, post_w, n_docs, n_terms, block_docs,
                                                       (int)n_blocks, q_ptr, q_terms, q_vals, k, cap, id_offset,
                                                       part_scores, part_ids); count_launch();
    B200_CUDA_CHECK(cudaGetLastError());
    return launch_merge(part_scores, part_ids, n_queries, nullptr, (int)(n_blocks * k), k, nullptr, out_scores, out_ids,
                        out_counts, st);
}

}  // extern "C"
