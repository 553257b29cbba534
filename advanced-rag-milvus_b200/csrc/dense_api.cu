// dense_api.cu -- b200rag_dense_topk: dispatch between the tensor-core scan (dense_tc.cu) and the exact
// CUDA-core scan (dense_exact.cu).
#include "common.cuh"

namespace b200rag {
size_t exact_workspace_bytes(int64_t n_rows, int dim, int n_q, int k);
int run_exact(const void* corpus16, int64_t n_rows, int dim, int dtype, const void* queries16, int n_launch,
              const int32_t* q_list, int k, int64_t id_offset, double* out_scores, int64_t* out_ids,
              void* workspace, size_t workspace_bytes, cudaStream_t st, const int32_t* n_active, int slot_base,
              const uint32_t* row_mask);
size_t tensor_workspace_bytes(int64_t n_rows, int dim, int n_q, int k);
bool tensor_supported(int64_t n_rows, int dim, int n_q, int k);
int run_tensor(const void* corpus16, int64_t n_rows, int dim, int dtype, const void* queries16, int n_q, int k,
               int64_t id_offset, double* out_scores, int64_t* out_ids, int32_t* out_flags, double row_norm_bound,
               float* out_err, int with_fallback, void* workspace, size_t workspace_bytes, cudaStream_t st,
               const uint32_t* row_mask);
void profile_next_scan(void* a, void* b);
}  // namespace b200rag

using namespace b200rag;

extern "C" {

int b200rag_profile_next_scan(void* start_event, void* stop_event) {
    profile_next_scan(start_event, stop_event);
    return B200RAG_OK;
}

size_t b200rag_dense_topk_workspace_bytes(int64_t n_rows, int32_t dim, int32_t n_queries, int32_t k, int32_t mode) {
    if (n_rows < 0 || dim <= 0 || n_queries < 0 || k <= 0) return 0;
    if (mode == B200RAG_DENSE_EXACT) return exact_workspace_bytes(n_rows, dim, n_queries, k);
    if (mode == B200RAG_DENSE_AUTO && !tensor_supported(n_rows, dim, n_queries, k)) return exact_workspace_bytes(n_rows, dim, n_queries, k);
    return tensor_workspace_bytes(n_rows, dim, n_queries, k);
}

int b200rag_dense_topk(const void* corpus16, int64_t n_rows, int32_t dim, int32_t dtype,
                       const void* queries16, int32_t n_queries, int32_t k, int64_t id_offset,
                       double* out_scores, int64_t* out_ids, int32_t* out_flags,
                       double row_norm_bound, float* out_err,
                       void* workspace, size_t workspace_bytes, int32_t mode, void* stream) {
    return b200rag_dense_topk_masked(corpus16, n_rows, dim, dtype, queries16, n_queries, k, id_offset, out_scores, out_ids,
                                     out_flags, row_norm_bound, out_err, nullptr, workspace, workspace_bytes, mode, stream);
}

int b200rag_dense_topk_masked(const void* corpus16, int64_t n_rows, int32_t dim, int32_t dtype,
                              const void* queries16, int32_t n_queries, int32_t k, int64_t id_offset,
                              double* out_scores, int64_t* out_ids, int32_t* out_flags,
                              double row_norm_bound, float* out_err, const uint32_t* row_mask,
                              void* workspace, size_t workspace_bytes, int32_t mode, void* stream) {
    B200_REQUIRE(n_queries >= 0, "dense_topk: bad n_queries=%d", n_queries);
    if (n_queries == 0) return B200RAG_OK;
    B200_REQUIRE(queries16 && out_scores && out_ids && workspace, "dense_topk: null pointer");
    B200_REQUIRE(corpus16 || n_rows == 0, "dense_topk: null corpus");
    B200_REQUIRE(n_rows >= 0 && n_rows < ((int64_t)1 << 40), "dense_topk: bad n_rows %lld", (long long)n_rows);
    B200_REQUIRE(dim > 0 && dim % 8 == 0 && dim <= 4096, "dense_topk: dim must be a multiple of 8 in (0,4096], got %d", dim);
    B200_REQUIRE(n_queries >= 0 && k > 0 && k <= 2048, "dense_topk: bad n_queries=%d / k=%d", n_queries, k);
    B200_REQUIRE(dtype == B200RAG_F16 || dtype == B200RAG_BF16, "dense_topk: bad dtype %d", dtype);
    B200_REQUIRE(((uintptr_t)corpus16 & 15) == 0 && ((uintptr_t)queries16 & 15) == 0 && ((uintptr_t)workspace & 255) == 0,
                 "dense_topk: corpus/queries must be 16-byte aligned and the workspace 256-byte aligned");
    if (n_queries == 0) return B200RAG_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (mode == B200RAG_DENSE_AUTO && !tensor_supported(n_rows, dim, n_queries, k)) mode = B200RAG_DENSE_EXACT;   // e.g. k > 1000
    if (mode == B200RAG_DENSE_EXACT) {
        if (out_flags) B200_CUDA_CHECK(cudaMemsetAsync(out_flags, 0, (size_t)n_queries * sizeof(int32_t), st));
        if (out_err) B200_CUDA_CHECK(cudaMemsetAsync(out_err, 0, (size_t)n_queries * sizeof(float), st));
        return run_exact(corpus16, n_rows, dim, dtype, queries16, n_queries, nullptr, k, id_offset, out_scores, out_ids,
                         workspace, workspace_bytes, st, nullptr, 0, row_mask);
    }
    if (mode == B200RAG_DENSE_AUTO || mode == B200RAG_DENSE_TENSOR || mode == B200RAG_DENSE_APPROX) {
        B200_REQUIRE(row_norm_bound > 0.0 && row_norm_bound < 1e30, "dense_topk: row_norm_bound must be positive (got %g)",
                     row_norm_bound);
        if (mode == B200RAG_DENSE_APPROX && !tensor_supported(n_rows, dim, n_queries, k)) {
            set_error("dense_topk: k=%d is beyond the tensor-core path (no approximate mode for it; use B200RAG_DENSE_AUTO)", k);
            return B200RAG_E_UNSUPPORTED;
        }
        return run_tensor(corpus16, n_rows, dim, dtype, queries16, n_queries, k, id_offset, out_scores, out_ids, out_flags,
                          row_norm_bound, out_err, mode == B200RAG_DENSE_AUTO ? 1 : (mode == B200RAG_DENSE_APPROX ? 2 : 0), workspace,
                          workspace_bytes, st, row_mask);
    }
    set_error("dense_topk: unknown mode %d", mode);
    return B200RAG_E_INVALID;
}

}  // extern "C"
