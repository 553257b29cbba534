// filter_mask.cu -- metadata predicate -> row bit mask on the GPU.
//
// Replaces the server-side evaluation of the `expr=filters` argument of Collection.search (reference
// src/advanced_rag/indexing.py:505-523); the expression strings come from HybridRetriever._build_filter_expression
// (reference retrieval.py:565-632): conjunctions of `field op value` over eight whitelisted scalar fields of the
// collection schema (indexing.py:191-225: VARCHAR doc_id / chunk_id / timestamp, INT64 chunk_index / token_count, FLOAT
// entropy / redundancy / domain_density).  The host mirror (b200rag/index_manager.py) keeps those fields as typed device
// columns; one launch of this kernel turns a parsed expression into the u32 bit mask the *_topk_masked entry points take.
// HBM-bound: 8 (or 4) bytes per row and term read, 1 bit per row written; one warp produces one 32-row word by ballot.
#include "common.cuh"

namespace b200rag {

constexpr int FM_MAX_TERMS = 16;
constexpr int FM_THREADS = 256;

struct FilterTerms {
    b200rag_filter_term t[FM_MAX_TERMS];
    int n;
};

template <typename T>
__device__ __forceinline__ bool fm_cmp(T a, T b, int op) {
    switch (op) {
        case B200RAG_OP_EQ: return a == b;
        case B200RAG_OP_NE: return a != b;
        case B200RAG_OP_GE: return a >= b;
        case B200RAG_OP_LE: return a <= b;
        case B200RAG_OP_GT: return a > b;
        default: return a < b;
    }
}

__global__ void __launch_bounds__(FM_THREADS)
filter_mask_kernel(const FilterTerms terms, int64_t n_rows, const uint32_t* __restrict__ and_mask, uint32_t* __restrict__ out_mask,
                   unsigned long long* __restrict__ out_count) {
    const int lane = threadIdx.x & 31;
    const int64_t n_words = (n_rows + 31) >> 5;
    const int64_t warps_total = (int64_t)gridDim.x * (FM_THREADS / 32);
    int local = 0;
    for (int64_t w = (int64_t)blockIdx.x * (FM_THREADS / 32) + (threadIdx.x >> 5); w < n_words; w += warps_total) {
        const int64_t row = (w << 5) + lane;
        bool ok = row < n_rows;
        for (int i = 0; ok && i < terms.n; ++i) {
            const b200rag_filter_term& t = terms.t[i];
            if (t.kind == B200RAG_COL_F64) {
                const double v = static_cast<const double*>(t.column)[row];
                ok = v == v && fm_cmp<double>(v, t.fvalue, t.op);                 // NaN = missing value: never matches
            } else if (t.kind == B200RAG_COL_I64) {
                const long long v = static_cast<const long long*>(t.column)[row];
                ok = v != INT64_MIN && fm_cmp<long long>(v, (long long)t.ivalue, t.op);
            } else if (t.kind == B200RAG_COL_I64_AS_F64) {
                const long long v = static_cast<const long long*>(t.column)[row];
                ok = v != INT64_MIN && fm_cmp<double>((double)v, t.fvalue, t.op);
            } else if (t.kind == B200RAG_COL_CODE) {
                const int c = static_cast<const int*>(t.column)[row];
                // code < 0 = missing value.  With a table: the host evaluated the operator on every dictionary entry; without:
                // == / != against the literal's code (ivalue; a literal that is not in the dictionary has code -2)
                ok = c >= 0 && (t.lut ? (c < t.lut_size && t.lut[c] != 0) : fm_cmp<long long>((long long)c, (long long)t.ivalue, t.op));
            } else {
                ok = false;                                                       // B200RAG_COL_NEVER: type mismatch
            }
        }
        unsigned word = __ballot_sync(0xffffffffu, ok);
        if (and_mask) word &= and_mask[w];
        if (lane == 0) {
            out_mask[w] = word;
            local += __popc(word);
        }
    }
    if (out_count && lane == 0 && local) atomicAdd(out_count, (unsigned long long)local);
}

}  // namespace b200rag

using namespace b200rag;

extern "C" {

int b200rag_filter_mask(const b200rag_filter_term* terms, int32_t n_terms, int64_t n_rows, const uint32_t* and_mask,
                        uint32_t* out_mask, int64_t* out_count, void* stream) {
    B200_REQUIRE(n_terms >= 0 && n_terms <= FM_MAX_TERMS, "filter_mask: at most %d terms, got %d", FM_MAX_TERMS, n_terms);
    B200_REQUIRE(n_rows >= 0 && (terms || n_terms == 0) && (out_mask || n_rows == 0), "filter_mask: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (out_count) B200_CUDA_CHECK(cudaMemsetAsync(out_count, 0, sizeof(int64_t), st));
    if (n_rows == 0) return B200RAG_OK;
    FilterTerms ft;
    ft.n = n_terms;
    for (int i = 0; i < n_terms; ++i) {
        ft.t[i] = terms[i];
        B200_REQUIRE(terms[i].op >= B200RAG_OP_EQ && terms[i].op <= B200RAG_OP_LT, "filter_mask: bad operator %d", terms[i].op);
        B200_REQUIRE(terms[i].kind >= B200RAG_COL_F64 && terms[i].kind <= B200RAG_COL_NEVER, "filter_mask: bad column kind %d", terms[i].kind);
        B200_REQUIRE(terms[i].kind == B200RAG_COL_NEVER || terms[i].column, "filter_mask: null column in term %d", i);
        B200_REQUIRE(terms[i].kind != B200RAG_COL_CODE || terms[i].lut || terms[i].op <= B200RAG_OP_NE,
                     "filter_mask: dictionary term %d needs a table for an ordering operator", i);
    }
    const int64_t n_words = (n_rows + 31) >> 5;
    int64_t blocks = (n_words + FM_THREADS / 32 - 1) / (FM_THREADS / 32);
    if (blocks > 148 * 16) blocks = 148 * 16;
    filter_mask_kernel<<<(unsigned)blocks, FM_THREADS, 0, st>>>(ft, n_rows, and_mask, out_mask,
                                                               reinterpret_cast<unsigned long long*>(out_count)); count_launch();
    B200_CUDA_CHECK(cudaGetLastError());
    return B200RAG_OK;
}

}  // extern "C"
