// finish.cuh -- parameters of the finish stage (K2) shared by its two kernel generations (dense_tc.cu, dense_finish.cu).
#pragma once
#include "tc_common.cuh"

namespace b200rag {

struct FinishParams {
    const uint16_t* corpus;
    const uint16_t* queries;
    int64_t n_rows;
    int dim;
    int n_q;
    int k;
    int kprime;
    int cap;
    int nqb;
    int n_chunks;
    int topk_cap;        // BlockTopK capacity
    int stage_rows;      // candidate rows staged per re-score batch (<= FN_THREADS / 8; fewer for very wide vectors)
    int64_t id_offset;
    double row_norm_bound;
    const unsigned long long* cand;
    const int* cand_cnt;
    const unsigned int* gthr;
    double* out_scores;
    int64_t* out_ids;
    int32_t* out_flags;
    int32_t* flag_list;   // compacted list of flagged queries
    int32_t* n_flagged;
    float* err_max;       // optional [n_q]: max |tensor score - exact score| over the re-scored candidates
    // tier-0 re-scan (dense_finish.cu only): launch slot -> original query, and the number of live slots (device scalar)
    const int32_t* q_list = nullptr;
    const int32_t* gate = nullptr;
    int approx = 0;       // B200RAG_DENSE_APPROX: rank by the tensor-core (fp32) scores, no fp64 re-score, no proof
    // split finish (launch_finish3): the selection of every query, handed from the select kernel to the wide re-score kernel
    // and on to the rank kernel through the workspace
    unsigned long long* fin_sel = nullptr;   // [n_q, kprime] (tensor score bits << 32) | row
    double* fin_exact = nullptr;             // [n_q, kprime] canonical fp64 scores
    int* fin_n = nullptr;                    // [n_q] candidates selected
    float* fin_m = nullptr;                  // [n_q] bound on the tensor score of every row outside the selection
};


// Second generation (dense_finish.cu): one 128-thread CTA per query, warp-level selection, one thread per candidate row.
size_t finish2_smem_bytes(int dim, int kprime);
int launch_finish2(const FinishParams& fp, int dtype, cudaStream_t st);
// Split finish: the same selection kernel, then the fp64 re-score as its own grid over (query, 16 candidate rows) -- every
// candidate row of the batch in flight at once instead of 16 per query -- then a rank / proof / emit kernel.
size_t finish3_workspace_bytes(int n_q, int kprime);
int launch_finish3(FinishParams fp, int dtype, void* fin_ws, cudaStream_t st);
// Warp-per-query replacement of sample_threshold_kernel.
int launch_sample_threshold2(const unsigned long long* cand, int cap, int nqb, int n_chunks, int rank, unsigned int* gthr,
                             cudaStream_t st, const int* gate);

}  // namespace b200rag
