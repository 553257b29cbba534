/*
 * b200rag.h -- C ABI of the B200-native hybrid-retrieval engine (libb200rag.so).
 *
 * This is the drop-in boundary for the retrieval hot path of rnaarla/advanced-rag-milvus.  The reference has
 * no FFI of its own: its seam is the duck-typed index manager that HybridRetriever calls
 * (reference src/advanced_rag/retrieval.py:113-131), whose one arithmetic entry point is
 * MilvusIndexManager.search() -> pymilvus Collection.search() (src/advanced_rag/indexing.py:445-551, call site
 * :505-523), followed in-process by _fuse_results (retrieval.py:421-491) and _mmr_diversify
 * (retrieval.py:493-516).  Each function below names the reference call it replaces.  INTEGRATION.md shows the
 * ctypes binding a maintainer of the reference adds.
 *
 * Conventions
 *   - plain C, no exceptions.  Every function returns 0 on success or a negative B200RAG_E_* code;
 *     b200rag_last_error() returns a thread-local message for the last failure on the calling thread.
 *   - all array arguments are DEVICE pointers owned by the caller (e.g. torch tensors' data_ptr()), contiguous,
 *     row major.  Nothing caller-visible is allocated by the library; scratch is a caller-provided workspace whose
 *     size the matching *_workspace_bytes() function reports.  Workspaces must be 256-byte aligned.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream).
 *   - 16-bit vector data is passed as raw bit patterns (fp16 or bf16, see b200rag_dtype).
 *   - ranking rule everywhere: score descending, then id ascending.  Unused output slots: id -1, score -inf.
 *   - there is NO CPU fallback: every function fails with B200RAG_E_CUDA if no sm_100 device is usable.
 *
 * Canonical arithmetic (identical, bit for bit, to oracle/exact_scan.c):
 *   dense   sum_d q[d]*x[d] over the stored 16-bit values, fp64, 8 interleaved lanes, fixed combine tree.
 *   cosine  = dense score of rows/queries normalised once by b200rag_prepare_rows(normalize=1).
 *   sparse  fp32 fmaf chain over the query's terms in ascending term id.
 *   rrf     fp64:  fused[id] += (1.0 / (rrf_k + rank)) * w   in list order; stable sort by score desc.
 *   mmr     fp64:  lambda*rel - (1-lambda)*max_jaccard, strict '>' so the earliest candidate wins ties.
 */
#ifndef B200RAG_H_
#define B200RAG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200RAG_ABI_VERSION 2

typedef enum { B200RAG_F16 = 0, B200RAG_BF16 = 1 } b200rag_dtype;

/* dense scan engine selection (b200rag_dense_topk `mode`) */
typedef enum {
    B200RAG_DENSE_AUTO = 0,   /* tensor-core scan + exact re-score, exact fallback for flagged queries */
    B200RAG_DENSE_EXACT = 1,  /* CUDA-core fp64 canonical scan only (slow; ground truth / fallback)    */
    B200RAG_DENSE_TENSOR = 2, /* tensor-core scan + exact re-score, flags reported, NO fallback        */
    B200RAG_DENSE_APPROX = 3  /* APPROXIMATE: ranks by the tensor-core fp32 scores (hardware accumulation order), no fp64 re-score,
                                 no completeness proof, no fallback.  out_scores = those fp32 scores.  Near-ties may swap places or
                                 cross rank k; bench.py reports recall@k against the exact modes */
} b200rag_dense_mode;

enum {
    B200RAG_OK = 0,
    B200RAG_E_INVALID = -1,    /* bad argument (null pointer, size, alignment, unsupported k / dim) */
    B200RAG_E_WORKSPACE = -2,  /* workspace too small or misaligned */
    B200RAG_E_CUDA = -3,       /* CUDA runtime / driver error, or no sm_100 device */
    B200RAG_E_UNSUPPORTED = -4
};

/* Thread-local description of the last error on this thread ("" if none). */
const char* b200rag_last_error(void);
int b200rag_abi_version(void);
/* Number of CUDA kernels this library has launched in this process so far (bench.py reports the difference over its
 * timed region as gpu_launches). */
uint64_t b200rag_kernel_launch_count(void);
/* Fills SM count and compute capability of the current device. */
int b200rag_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------------------------------------------------
 * Row preparation (ingest side and query side).  Replaces what Milvus does at insert time for a COSINE index
 * (reference indexing.py:143-180 creates the collections with metric_type COSINE; vectors are inserted as fp32
 * at :372,426) and the `query_embedding.tolist()` marshalling at indexing.py:500.
 *   in_f32 [n_rows, dim] fp32  ->  out16 [n_rows, dim] 16-bit patterns of `dtype`
 *   normalize=1: canonical L2 normalisation (fp64 sequential sum of squares, one rounding); 0: plain RNE cast.
 */
int b200rag_prepare_rows(const float* in_f32, void* out16, int64_t n_rows, int32_t dim, int32_t dtype,
                         int32_t normalize, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Dense exact top-k  (K1 + K2).  Replaces Collection.search on "semantic_index" / "domain_index"
 * (reference indexing.py:505-523, reached from retrieval.py:341-365 and :397-419) for a whole batch of queries.
 *   corpus16  [n_rows, dim]   stored rows of this shard (dim % 8 == 0, 16-byte aligned base)
 *   queries16 [n_queries, dim]
 *   out_scores f64 [n_queries, k]  canonical scores;  out_ids i64 [n_queries, k] = local row + id_offset
 *   out_flags  i32 [n_queries] or NULL: bit1 = the query went through the widened tier-0 re-scan (k' = 640 candidates, for
 *              tie groups across the edge of the first pass' candidate set); bit0 = the tensor-core candidate set could not be PROVEN complete for
 *              this query (near-ties beyond the slack); in AUTO mode such queries were re-run on the exact path,
 *              so results are always exact and the flag is informational.  AUTO never synchronises with the host: the
 *              fallback is launched unconditionally and gated on the device by the number of flagged queries.  Shapes
 *              the tensor-core path cannot serve (k > ~1000) are routed to the exact scan by AUTO on its own.
 *   row_norm_bound  upper bound on the L2 norm of any stored row (1.001 for rows prepared with normalize=1); it
 *              sizes the error margin of the completeness proof of the tensor-core path.  Ignored in EXACT mode.
 *   out_err    f32 [n_queries] or NULL: max |tensor-core score - canonical score| over the re-scored candidates
 *              (diagnostic; lets callers verify the error bound the proof relies on).
 */
size_t b200rag_dense_topk_workspace_bytes(int64_t n_rows, int32_t dim, int32_t n_queries, int32_t k, int32_t mode);
int b200rag_dense_topk(const void* corpus16, int64_t n_rows, int32_t dim, int32_t dtype,
                       const void* queries16, int32_t n_queries, int32_t k, int64_t id_offset,
                       double* out_scores, int64_t* out_ids, int32_t* out_flags,
                       double row_norm_bound, float* out_err,
                       void* workspace, size_t workspace_bytes, int32_t mode, void* stream);

/* The same search restricted to the rows a metadata filter allows (reference: the `expr=filters` argument of
 * Collection.search, indexing.py:505-523, built by HybridRetriever._build_filter_expression, retrieval.py:565-632).
 *   row_mask  u32 [ceil(n_rows / 32)]: bit (row & 31) of word (row >> 5) set = row allowed; NULL = no filter.
 * The mask is applied inside the kernels (sample pass, survivors of the scan epilogue, exact scan); results are the exact
 * top-k of the allowed rows, fewer than k hits are padded with id -1 / score -inf. */
int b200rag_dense_topk_masked(const void* corpus16, int64_t n_rows, int32_t dim, int32_t dtype,
                              const void* queries16, int32_t n_queries, int32_t k, int64_t id_offset,
                              double* out_scores, int64_t* out_ids, int32_t* out_flags,
                              double row_norm_bound, float* out_err, const uint32_t* row_mask,
                              void* workspace, size_t workspace_bytes, int32_t mode, void* stream);

/* Profiling hook (not part of the data path): when both handles are non-NULL cudaEvent_t values, the NEXT
 * tensor-core b200rag_dense_topk call on this thread records `start` immediately before and `stop` immediately
 * after its scan kernel on the call's stream, then clears the hook.  bench.py uses it to time the dominant kernel
 * inside the timed region without a profiler. */
int b200rag_profile_next_scan(void* start_event, void* stop_event);
/* Debug/profiling: per-CTA cycle counters, written into a DEVICE buffer the caller owns (the library keeps no global
 * state: the registration is per calling thread and holds until it is cleared with a NULL buffer).
 *   kind 0 = tensor-core dense scan: 16 u64 slots per CTA, 256 CTAs (MMA total, MMA wait-for-data, MMA wait-for-epilogue,
 *            MMA wait-for-queries, producer total, producer wait-for-slot, epilogue total, epilogue wait-for-accumulator,
 *            epilogue compaction, epilogue query load, #compactions, #slow-path groups); needs n_slots >= 4096.
 *   kind 1 = sparse scan: 12 u64 slots per CTA, first 1024 CTAs (see sparse_bm25.cu); needs n_slots >= 12288.
 * A scan launched from this thread while a large-enough buffer is registered zeroes it and fills it. */
int b200rag_debug_set_stats_buffer(int32_t kind, uint64_t* device_buf, size_t n_slots);

/* A/B options for measurements (tools/scan_ab.py, tools/sparse_ab.py).  Process-wide integer knobs, set explicitly and read
 * with atomic loads -- nothing is read from the environment.  value < 0 restores the built-in default.
 *   "scan_version" 1|3, "qg_span", "epi" 0|1, "no_sample" 0|1, "stage_rows" 8|16|32, "sample_mult", "mmr_path" (0 auto, 1 general
 *   kernel), "sparse_slices", "sparse_flags", "no_tier0" 0|1, "finish_version" 0 auto | 1 | 2 | 3.
 * b200rag_get_option returns the stored value (-1 = default in force, -2 = unknown name). */
int b200rag_set_option(const char* name, int64_t value);
int64_t b200rag_get_option(const char* name);

/* ---------------------------------------------------------------------------------------------------------
 * Metadata predicate -> row bit mask.  Replaces the server-side evaluation of `expr=filters` in Collection.search
 * (reference indexing.py:505-523); expressions are the conjunctions HybridRetriever._build_filter_expression emits
 * (reference retrieval.py:565-632) over the scalar fields of the collection schema (indexing.py:191-225).
 *   terms    HOST array of n_terms (<= 16) parsed terms; `column` / `lut` are DEVICE pointers with one entry per row /
 *            per dictionary code.  Missing values never match: NaN (f64), INT64_MIN (i64), code < 0 (dictionary).
 *   and_mask optional u32 [ceil(n_rows/32)] ANDed into the result (e.g. the live-row mask after deletes); may be NULL
 *   out_mask u32 [ceil(n_rows/32)], bit (row & 31) of word (row >> 5) = row satisfies every term
 *   out_count optional DEVICE i64: number of set bits */
typedef enum { B200RAG_OP_EQ = 0, B200RAG_OP_NE = 1, B200RAG_OP_GE = 2, B200RAG_OP_LE = 3, B200RAG_OP_GT = 4, B200RAG_OP_LT = 5 } b200rag_filter_op;
typedef enum {
    B200RAG_COL_F64 = 0,         /* double column compared with fvalue */
    B200RAG_COL_I64 = 1,         /* int64 column compared with ivalue */
    B200RAG_COL_I64_AS_F64 = 2,  /* int64 column compared with fvalue (non-integral literal) */
    B200RAG_COL_CODE = 3,        /* int32 dictionary codes: row matches iff lut[code] != 0 (the host evaluated op on the dictionary);
                                    lut == NULL: op must be == or != and is applied to (code, ivalue) */
    B200RAG_COL_NEVER = 4        /* literal type does not fit the column: no row matches */
} b200rag_filter_kind;
typedef struct {
    const void* column;
    const uint8_t* lut;
    double fvalue;
    int64_t ivalue;
    int32_t kind;
    int32_t op;
    int32_t lut_size;
    int32_t reserved;
} b200rag_filter_term;
int b200rag_filter_mask(const b200rag_filter_term* terms, int32_t n_terms, int64_t n_rows, const uint32_t* and_mask,
                        uint32_t* out_mask, int64_t* out_count, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Sparse inner-product top-k over doc-range-blocked postings (K3).  Replaces Collection.search on
 * "sparse_index" (reference indexing.py:472,487-498,505-523, reached from retrieval.py:367-395).
 * Postings layout ("blocked CSR"): documents are cut into blocks of `block_docs` consecutive rows
 * (block_docs <= 32768; 16384 lets two CTAs share an SM); inside block b the postings of term t occupy
 *   [blk_term_ptr[b*(n_terms+1)+t], blk_term_ptr[b*(n_terms+1)+t+1])  of post_doc / post_w,
 * post_doc holding the row index RELATIVE to the block start (u16), ascending.
 * Queries are a CSR over terms: q_ptr i64 [n_queries+1], q_terms i32 ascending per query, q_vals f32.
 * Only documents sharing at least one term with the query are candidates; out_counts[q] = hits returned.
 */
size_t b200rag_sparse_topk_workspace_bytes(int64_t n_docs, int32_t block_docs, int32_t n_queries, int32_t k);
int b200rag_sparse_topk(const int64_t* blk_term_ptr, const uint16_t* post_doc, const float* post_w,
                        int64_t n_docs, int32_t n_terms, int32_t block_docs,
                        const int64_t* q_ptr, const int32_t* q_terms, const float* q_vals,
                        int32_t n_queries, int32_t k, int64_t id_offset,
                        float* out_scores, int64_t* out_ids, int32_t* out_counts,
                        void* workspace, size_t workspace_bytes, void* stream);

/* Filtered variant: doc_mask u32 [ceil(n_docs / 32)] as in b200rag_dense_topk_masked (NULL = no filter). */
int b200rag_sparse_topk_masked(const int64_t* blk_term_ptr, const uint16_t* post_doc, const float* post_w,
                               int64_t n_docs, int32_t n_terms, int32_t block_docs,
                               const int64_t* q_ptr, const int32_t* q_terms, const float* q_vals,
                               int32_t n_queries, int32_t k, int64_t id_offset,
                               float* out_scores, int64_t* out_ids, int32_t* out_counts, const uint32_t* doc_mask,
                               void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * k-way merge of gathered candidate lists (K4): the reduce after the NCCL all-gather of per-GPU top-k.  Stands
 * in for the reduce Milvus' proxy performs across its shards (reference indexing.py:91,234-239 num_shards).
 *   cand_scores f64 [n_queries, n_cand], cand_ids i64 [n_queries, n_cand]  (id < 0 = empty slot)
 */
size_t b200rag_merge_topk_workspace_bytes(int32_t n_queries, int32_t n_cand, int32_t k);
int b200rag_merge_topk(const double* cand_scores, const int64_t* cand_ids, int32_t n_queries, int32_t n_cand,
                       int32_t k, double* out_scores, int64_t* out_ids,
                       void* workspace, size_t workspace_bytes, void* stream);

/* Same reduce, reading the NCCL all-gather buffer in place: gathered i64 [n_ranks, 2, n_queries, k] = per rank a plane of
 * k fp64 score bit patterns per query followed by a plane of k ids per query (id < 0 = empty slot) -- i.e. the out_scores and
 * out_ids arrays of b200rag_dense_topk, which a rank points straight at the two halves of its send buffer.  No pack, unpack
 * or transpose pass between the search, the collective and the merge.  out_counts i32 [n_queries] (may be NULL): valid results
 * per query. */
int b200rag_merge_gathered(const int64_t* gathered, int32_t n_ranks, int32_t n_queries, int32_t k,
                           double* out_scores, int64_t* out_ids, int32_t* out_counts, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Weighted Reciprocal Rank Fusion (K5).  Replaces HybridRetriever._fuse_results (reference
 * retrieval.py:421-491) for a batch.
 *   list_ids  i64 [n_lists, n_queries, k_max]  ranked ids, list order = semantic, sparse[, domain]
 *   list_len  i32 [n_lists, n_queries]
 *   weights   f64 [n_queries, n_lists]         (dense_weight, sparse_weight[, 0.2]) per query
 *   out_ids   i64 [n_queries, n_lists*k_max]   fused order (score desc, ties in first-seen order)
 *   out_scores f64 same shape; out_mask i32 same shape (bit i: list i held the id);
 *   out_first i32 same shape: position (list*k_max + rank0) of the hit that supplies the payload
 *             (the reference keeps the first list's dict, retrieval.py:441,449-450,460-461);
 *   out_n     i32 [n_queries] number of fused ids.
 */
size_t b200rag_rrf_fuse_workspace_bytes(int32_t n_lists, int32_t n_queries, int32_t k_max);
int b200rag_rrf_fuse(const int64_t* list_ids, const int32_t* list_len, int32_t n_lists, int32_t n_queries,
                     int32_t k_max, const double* weights, int32_t rrf_k,
                     int64_t* out_ids, double* out_scores, int32_t* out_mask, int32_t* out_first, int32_t* out_n,
                     void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * The tail of the fusion stage for a batch: which fused entries a query returns and their columns.  Replaces the list
 * slicing and per-hit tagging at the end of HybridRetriever._retrieve_inner (reference retrieval.py:322-333, 441-461,
 * 485-491, 512-516): slot j of query q is fused position picks[q][j] (clamped at 0) if use_mmr[q], else j; valid while
 * j < min(fused_n[q], top_k[q]).
 *   fused_*      the outputs of b200rag_rrf_fuse, [n_queries, tot] with tot = n_lists * k_max
 *   picks        i32 [n_queries, t_max] from b200rag_mmr_select, or NULL;  use_mmr i32 [n_queries] or NULL
 *   list_scores  f64 [n_lists, n_queries, k_max]  the retrieval methods' own scores (original_score of a hit)
 *   out_rows i64 [n_queries, t_max] (-1 pads), out_scores f64 (-inf pads), out_mask i32 (0 pads),
 *   out_first_method i32 (index of the list whose hit supplies the payload), out_original f64, out_n i32 [n_queries] */
int b200rag_fuse_select(const int64_t* fused_ids, const double* fused_scores, const int32_t* fused_mask, const int32_t* fused_first,
                        const int32_t* fused_n, int32_t n_queries, int32_t tot, const int32_t* picks, const int32_t* use_mmr,
                        const int32_t* top_k, const double* list_scores, int32_t n_lists, int32_t k_max, int32_t t_max,
                        int64_t* out_rows, double* out_scores, int32_t* out_mask, int32_t* out_first_method, double* out_original,
                        int32_t* out_n, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Greedy MMR selection on token-set Jaccard (K6).  Replaces HybridRetriever._mmr_diversify (reference
 * retrieval.py:493-516) for a batch.
 *   cand_doc  i32 [n_queries, n_max]  row of each candidate in the token CSR, in fused order
 *   cand_rel  f64 [n_queries, n_max]  fused scores;  cand_n i32 [n_queries]
 *   doc_tok_ptr i64 [n_docs+1], doc_tok_ids i32: per document the SORTED UNIQUE token ids of
 *             set(content.lower().split()) (retrieval.py:497); ids < vocab_size
 *   lambda f64 [n_queries], k_sel i32 [n_queries] (per-query profile values, retrieval.py:142-213)
 *   out_pick i32 [n_queries, k_max]  selected candidate positions in pick order; out_n i32 [n_queries]
 */
size_t b200rag_mmr_select_workspace_bytes(int32_t n_queries, int32_t n_max, int32_t vocab_size);
int b200rag_mmr_select(const int32_t* cand_doc, const double* cand_rel, const int32_t* cand_n, int32_t n_queries,
                       int32_t n_max, const int64_t* doc_tok_ptr, const int32_t* doc_tok_ids, int32_t vocab_size,
                       const double* lambda, const int32_t* k_sel, int32_t k_max,
                       int32_t* out_pick, int32_t* out_n,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------
 * Learned re-rank of a fused batch.  Replaces HybridRetriever.rerank with the LearnedRanker (reference retrieval.py:518-563,
 * ranker.py:109-125):  s = base_weight*score + method_bonus*popcount(method_mask) + recency_weight*recency  (fp64, the
 * reference's operation order), stable descending sort, first k_out.
 *   scores f64 [n_queries, t_max] fused scores, method_mask i32 same shape (bit l: list l returned the hit, as b200rag_rrf_fuse
 *   writes it), recency f64 same shape or NULL (= 0), n_in i32 [n_queries] valid entries per query
 *   out_pos i32 [n_queries, k_out] positions into the input (-1 padded), out_scores f64 same shape, out_n i32 [n_queries] */
int b200rag_rerank_learned(const double* scores, const int32_t* method_mask, const double* recency, const int32_t* n_in,
                           int32_t n_queries, int32_t t_max, double base_weight, double method_bonus, double recency_weight,
                           int32_t k_out, int32_t* out_pos, double* out_scores, int32_t* out_n, void* stream);

/* Mean pairwise token-set Jaccard of every query's result list.  Replaces RAGEvaluator._calculate_pairwise_similarity
 * (reference evaluation.py:327-344; diversity = 1 - it, :315-325): pairs i < j whose two token sets are non-empty, numpy's mean.
 *   docs i32 [n_queries, n_max] rows of the results in the token CSR (see b200rag_mmr_select), n_in i32 [n_queries]
 *   out_mean f64 [n_queries] (0 when no pair qualifies), out_pairs i32 [n_queries] number of pairs averaged */
size_t b200rag_pairwise_jaccard_workspace_bytes(int32_t n_queries, int32_t n_max);
int b200rag_pairwise_jaccard(const int32_t* docs, const int32_t* n_in, int32_t n_queries, int32_t n_max,
                             const int64_t* doc_tok_ptr, const int32_t* doc_tok_ids, double* out_mean, int32_t* out_pairs,
                             void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200RAG_H_ */
