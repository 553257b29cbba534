// sparse_bm25.cu -- sparse inner-product (BM25-weighted) top-k over doc-range-blocked postings  (K3).
//
// One CTA (512 threads) owns one QUERY and walks the document blocks of its slice one after the other:
//   * the block's per-document accumulators live in shared memory (fp32, block_docs <= 32768 -> 128 KB) and are zeroed
//     once per CTA; collecting a block's candidates puts them back to zero;
//   * the postings of up to 8 query terms inside the block are fetched TOGETHER (one coalesced u16 doc + f32 weight per
//     thread and term, a single exposed memory latency per block instead of one per term) and then applied in ascending
//     term id with a barrier between terms, so every document sees  acc = fmaf(qv, w, acc)  in the canonical order
//     (bit-identical to oracle/exact_scan.c:orc_sparse_topk).  Postings of one term hit distinct documents: no atomics
//     on the accumulators; a touched-bitmap marks the candidates;
//   * candidates are collected from the bitmap (a thread owns one or two 32-document words), losers against the running
//     k-th best are dropped on the spot, survivors go to the block-level streaming top-k (select.cuh), which lives for
//     the whole walk, so later blocks are filtered by the threshold earlier blocks established.
// Algorithmic HBM traffic = 6 bytes per posting of the query's terms.  grid = (queries, slices): with fewer queries
// than SMs the blocks are split into slices; merge_topk_kernel reduces the slices.
// (Round 1 first ran one CTA per (query, block) that re-scanned all 32768 documents of the block: 4.1 ms for 256
// queries over 1M documents = 0.8 % of the HBM roofline; see profiles/r1_hybrid_c4.md for this version.)
#include "common.cuh"
#include "select.cuh"

namespace b200rag {

int launch_merge(const double* cand_scores, const int64_t* cand_ids, int n_launch, const int32_t* q_list, int n_cand, int k,
                 double* out_scores_f64, float* out_scores_f32, int64_t* out_ids, int32_t* out_counts, cudaStream_t st);

constexpr int SP_THREADS = 512;     // with 16384-document blocks two CTAs fit per SM and overlap each other's latencies
constexpr int SP_TG = 8;       // query terms fetched together (one register pair per term and thread)

__global__ void __launch_bounds__(SP_THREADS, 2)
sparse_query_kernel(const int64_t* __restrict__ blk_term_ptr, const uint16_t* __restrict__ post_doc,
                    const float* __restrict__ post_w, int64_t n_docs, int n_terms, int block_docs, int n_blocks, int n_slices,
                    const int64_t* __restrict__ q_ptr, const int32_t* __restrict__ q_terms, const float* __restrict__ q_vals,
                    int k, int cap, int64_t id_offset, double* __restrict__ part_scores, int64_t* __restrict__ part_ids,
                    const uint32_t* __restrict__ doc_mask) {
    extern __shared__ __align__(16) char smem[];
    const int tid = threadIdx.x;
    const int q = blockIdx.x;
    const int slice = blockIdx.y;
    const int n_words = block_docs / 32;                                            // <= 2048
    float* acc = reinterpret_cast<float*>(smem);                                    // [block_docs]
    uint32_t* touched = reinterpret_cast<uint32_t*>(smem + (size_t)block_docs * 4);  // [n_words]
    char* p = smem + (size_t)block_docs * 4 + (size_t)n_words * 4;
    p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(p) + 15) & ~uintptr_t(15));
    __shared__ long long s_beg[SP_TG], s_end[SP_TG];
    __shared__ float s_qv[SP_TG];
    BlockTopK<SP_THREADS, uint32_t> tk;
    tk.attach(p, cap, k, SP_THREADS, /*start_digit=*/BlockTopK<SP_THREADS, uint32_t>::NLO + 3);
    tk.init();
    for (int i = tid; i < block_docs; i += SP_THREADS) acc[i] = 0.0f;
    for (int i = tid; i < n_words; i += SP_THREADS) touched[i] = 0u;
    __syncthreads();

    const int64_t qs = q_ptr[q];
    const int nq = (int)(q_ptr[q + 1] - qs);
    const int b0 = (int)((int64_t)slice * n_blocks / n_slices), b1 = (int)((int64_t)(slice + 1) * n_blocks / n_slices);
    // ranges of the first term group of the NEXT block are fetched one block ahead (pointer -> postings is a dependent chain)
    long long pre_s = 0, pre_e = 0;
    float pre_qv = 0.f;
    int my_t = -1;
    if (tid < SP_TG && tid < nq) {
        const int t = q_terms[qs + tid];
        if (t >= 0 && t < n_terms) { my_t = t; pre_qv = q_vals[qs + tid]; }
    }
    if (my_t >= 0 && b0 < b1) {
        const int64_t* tp0 = blk_term_ptr + (size_t)b0 * (n_terms + 1);
        pre_s = tp0[my_t]; pre_e = tp0[my_t + 1];
    }
    for (int blk = b0; blk < b1; ++blk) {
        const int64_t doc0 = (int64_t)blk * block_docs;
        const int64_t* tp = blk_term_ptr + (size_t)blk * (n_terms + 1);
        // ---- accumulate, SP_TG terms at a time ---------------------------------------------------------------
        for (int g0 = 0; g0 < nq; g0 += SP_TG) {
            if (tid < SP_TG) {
                long long s = 0, e = 0;
                float qv = 0.f;
                if (g0 == 0) {
                    s = pre_s; e = pre_e; qv = pre_qv;
                    if (my_t >= 0 && blk + 1 < b1) {                  // issue the next block's pointer loads now
                        const int64_t* tpn = tp + (n_terms + 1);
                        pre_s = tpn[my_t]; pre_e = tpn[my_t + 1];
                    }
                } else if (g0 + tid < nq) {
                    const int t = q_terms[qs + g0 + tid];
                    if (t >= 0 && t < n_terms) { s = tp[t]; e = tp[t + 1]; qv = q_vals[qs + g0 + tid]; }
                }
                s_beg[tid] = s; s_end[tid] = e; s_qv[tid] = qv;
            }
            __syncthreads();
            int dreg[SP_TG];
            float wreg[SP_TG];
#pragma unroll
            for (int j = 0; j < SP_TG; ++j) {
                const long long i = s_beg[j] + tid;
                dreg[j] = -1;
                wreg[j] = 0.f;
                if (i < s_end[j]) { dreg[j] = post_doc[i]; wreg[j] = post_w[i]; }
            }
#pragma unroll
            for (int j = 0; j < SP_TG; ++j) {
                const float qv = s_qv[j];
                if (dreg[j] >= 0) {
                    const int d = dreg[j];
                    acc[d] = fmaf(qv, wreg[j], acc[d]);
                    atomicOr(&touched[d >> 5], 1u << (d & 31));
                }
                for (long long i = s_beg[j] + tid + SP_THREADS; i < s_end[j]; i += SP_THREADS) {     // long posting lists
                    const int d = post_doc[i];
                    acc[d] = fmaf(qv, post_w[i], acc[d]);
                    atomicOr(&touched[d >> 5], 1u << (d & 31));
                }
                __syncthreads();
            }
        }
        // ---- collect: a thread owns words tid and tid + SP_THREADS of the bitmap -----------------------------------
        // Losers are dropped with ONE float compare against the running k-th best score (read once per block; the exact
        // (score, id) comparison happens only for the few candidates at or above it).
        unsigned long long m = 0;
        if (tid < n_words) { m = touched[tid]; touched[tid] = 0u; }
        if (tid + SP_THREADS < n_words) { m |= (unsigned long long)touched[tid + SP_THREADS] << 32; touched[tid + SP_THREADS] = 0u; }
        if (doc_mask && m) {
            // metadata filter: documents that are not allowed are dropped here (their accumulators still have to go back to
            // zero).  block_docs is a multiple of 32, so a bitmap word of the block is a word of the mask.
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
        float thr_f = tk.threshold_hi32_as_float();
        while (__syncthreads_or(m != 0ull)) {
            bool have = false;
            uint64_t h = 0;
            uint32_t l = 0;
            while (m) {
                const uint32_t lo32 = (uint32_t)m;
                const int bpos = lo32 ? __ffs((int)lo32) - 1 : 32 + __ffs((int)(uint32_t)(m >> 32)) - 1;
                m &= m - 1;
                const int d = bpos < 32 ? tid * 32 + bpos : (tid + SP_THREADS) * 32 + (bpos - 32);
                const float sc = acc[d];
                acc[d] = 0.0f;
                if (sc < thr_f) continue;
                h = (uint64_t)mono32(sc);
                l = ~(uint32_t)(doc0 + d);
                if (tk.passes(h, l)) { have = true; break; }
            }
            tk.offer(have, h, l);
            tk.settle();
            thr_f = tk.threshold_hi32_as_float();
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
    size_t smem = sparse_smem(block_docs, k, &cap);
    if (smem > 225 * 1024) {
        set_error("sparse_topk: block_docs=%d with k=%d needs %zu bytes of shared memory (max 230400); use a smaller block",
                  block_docs, k, smem);
        return B200RAG_E_UNSUPPORTED;
    }
    int64_t n_blocks = (n_docs + block_docs - 1) / block_docs;
    if (n_blocks < 1) n_blocks = 1;
    B200_REQUIRE(n_queries <= 2147483647 && n_blocks <= 2147483647, "sparse_topk: too many blocks");
    Workspace ws(workspace, workspace_bytes);
    double* part_scores = ws.take<double>((size_t)n_queries * n_blocks * k);
    int64_t* part_ids = ws.take<int64_t>((size_t)n_queries * n_blocks * k);
    if (!ws.ok()) {
        set_error("sparse_topk: workspace too small (%zu < %zu)", workspace_bytes, ws.off);
        return B200RAG_E_WORKSPACE;
    }
    B200_CUDA_CHECK(cudaFuncSetAttribute(sparse_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // one CTA per query; with fewer queries than ~2 per SM the blocks are cut into slices to fill the machine
    int sm_count = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    int64_t n_slices = (2 * (int64_t)sm_count) / n_queries;
    if (n_slices < 1) n_slices = 1;
    if (n_slices > n_blocks) n_slices = n_blocks;
    B200_REQUIRE(n_slices <= 65535, "sparse_topk: too many slices");
    dim3 grid((unsigned)n_queries, (unsigned)n_slices);
    sparse_query_kernel<<<grid, SP_THREADS, smem, st>>>(blk_term_ptr, post_doc, post_w, n_docs, n_terms, block_docs,
                                                       (int)n_blocks, (int)n_slices, q_ptr, q_terms, q_vals, k, cap, id_offset,
                                                       part_scores, part_ids, doc_mask); count_launch();
    B200_CUDA_CHECK(cudaGetLastError());
    return launch_merge(part_scores, part_ids, n_queries, nullptr, (int)(n_slices * k), k, nullptr, out_scores, out_ids,
                        out_counts, st);
}

}  // extern "C"
