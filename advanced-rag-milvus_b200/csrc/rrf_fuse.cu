// rrf_fuse.cu -- weighted Reciprocal Rank Fusion for a batch of queries  (K5).
//
// Restates HybridRetriever._fuse_results (reference src/advanced_rag/retrieval.py:421-491) on integer ids:
//   for each list in order (semantic, sparse, domain), rank r = 1..:  fused[id] += (1.0 / (rrf_k + r)) * w_list
//   result sorted by fused score descending with a STABLE sort => ties keep first-seen order.
// All arithmetic is fp64 with explicit round-to-nearest division / multiply / add in the reference's order, so the
// scores are bit-identical to Python's.
//
// One CTA per query.  Lists are processed one after the other (barrier in between) so that an id's contributions
// are added in list order; inside a list every thread owns one rank and inserts into an open-addressing hash table
// in shared memory.  An id that occurs twice inside ONE list (never produced by a search, but legal input) makes
// the addition order inside that list matter; the CTA detects it and replays that query serially on thread 0.
#include "common.cuh"
#include "select.cuh"

namespace b200rag {

constexpr int RRF_THREADS = 256;
constexpr long long RRF_EMPTY = -0x7fffffffffffffffLL - 1;

struct RrfSlot {
    long long id;
    double score;
    int first;     // list*k_max + rank0 of the first hit
    int mask;      // bit l: list l held the id
    int last_list; // last list that touched the slot (duplicate detection)
    int pad;
};

__device__ __forceinline__ uint32_t rrf_hash(long long id, uint32_t table_size) {
    unsigned long long x = (unsigned long long)id * 0x9E3779B97F4A7C15ull;
    return (uint32_t)(x >> 32) % table_size;
}
__device__ __forceinline__ uint32_t rrf_next(uint32_t h, uint32_t table_size) { return h + 1 == table_size ? 0u : h + 1; }

__global__ void __launch_bounds__(RRF_THREADS)
rrf_fuse_kernel(const int64_t* __restrict__ list_ids, const int32_t* __restrict__ list_len, int n_lists, int n_q, int k_max,
                const double* __restrict__ weights, int rrf_k, int table_size, int sort_cap,
                int64_t* __restrict__ out_ids, double* __restrict__ out_scores, int32_t* __restrict__ out_mask,
                int32_t* __restrict__ out_first, int32_t* __restrict__ out_n) {
    extern __shared__ __align__(16) char smem[];
    const int tid = threadIdx.x;
    const int q = blockIdx.x;
    RrfSlot* table = reinterpret_cast<RrfSlot*>(smem);
    uint64_t* key_hi = reinterpret_cast<uint64_t*>(smem + (size_t)table_size * sizeof(RrfSlot));
    uint32_t* key_lo = reinterpret_cast<uint32_t*>(key_hi + sort_cap);
    __shared__ int s_unique;
    __shared__ int s_dup;
    const uint32_t hmask = (uint32_t)table_size;   // table size (modulo hashing)

    for (int i = tid; i < table_size; i += RRF_THREADS) {
        table[i].id = RRF_EMPTY;
        table[i].score = 0.0;
        table[i].first = 0x7fffffff;
        table[i].mask = 0;
        table[i].last_list = -1;
    }
    if (tid == 0) { s_unique = 0; s_dup = 0; }
    __syncthreads();

    // ---- parallel path: one list at a time, one thread per rank ------------------------------------
    for (int l = 0; l < n_lists; ++l) {
        const int len = min(list_len[(size_t)l * n_q + q], k_max);
        const double w = weights[(size_t)q * n_lists + l];
        const int64_t* ids = list_ids + ((size_t)l * n_q + q) * k_max;
        for (int r = tid; r < len; r += RRF_THREADS) {
            const long long id = ids[r];
            const double contrib = __dmul_rn(__ddiv_rn(1.0, (double)(rrf_k + r + 1)), w);
            uint32_t h = rrf_hash(id, hmask);
            for (;;) {
                long long prev = atomicCAS(reinterpret_cast<unsigned long long*>(&table[h].id),
                                           (unsigned long long)RRF_EMPTY, (unsigned long long)id);
                if (prev == RRF_EMPTY || prev == id) break;
                h = rrf_next(h, hmask);
            }
            // within one list ids are unique (else s_dup is raised and the serial path redoes the query)
            int prev_list = atomicExch(&table[h].last_list, l);
            if (prev_list == l) {
                s_dup = 1;
            } else {
                table[h].score = __dadd_rn(table[h].score, contrib);
                table[h].mask |= 1 << l;
                if (table[h].first == 0x7fffffff) table[h].first = l * k_max + r;
            }
        }
        __syncthreads();
    }

    // ---- serial replay when a list repeated an id (addition order inside the list matters) ---------
    if (s_dup) {
        for (int i = tid; i < table_size; i += RRF_THREADS) {
            table[i].id = RRF_EMPTY;
            table[i].score = 0.0;
            table[i].first = 0x7fffffff;
            table[i].mask = 0;
        }
        __syncthreads();
        if (tid == 0) {
            for (int l = 0; l < n_lists; ++l) {
                const int len = min(list_len[(size_t)l * n_q + q], k_max);
                const double w = weights[(size_t)q * n_lists + l];
                const int64_t* ids = list_ids + ((size_t)l * n_q + q) * k_max;
                for (int r = 0; r < len; ++r) {
                    const long long id = ids[r];
                    const double contrib = __dmul_rn(__ddiv_rn(1.0, (double)(rrf_k + r + 1)), w);
                    uint32_t h = rrf_hash(id, hmask);
                    while (table[h].id != RRF_EMPTY && table[h].id != id) h = rrf_next(h, hmask);
                    table[h].id = id;
                    table[h].score = __dadd_rn(table[h].score, contrib);
                    table[h].mask |= 1 << l;
                    if (table[h].first == 0x7fffffff) table[h].first = l * k_max + r;
                }
            }
        }
        __syncthreads();
    }

    // ---- gather unique ids, sort by (score desc, first-seen asc) ------------------------------------
    for (int i = tid; i < table_size; i += RRF_THREADS) {
        if (table[i].id != RRF_EMPTY) {
            int j = atomicAdd(&s_unique, 1);
            key_hi[j] = mono64(table[i].score);
            key_lo[j] = ~(uint32_t)table[i].first;
        }
    }
    __syncthreads();
    const int n = s_unique;
    const int n2 = next_pow2_int(n > 1 ? n : 1);
    for (int i = n + tid; i < n2; i += RRF_THREADS) { key_hi[i] = 0; key_lo[i] = 0; }
    __syncthreads();
    for (int size = 2; size <= n2; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = tid; t < n2 / 2; t += RRF_THREADS) {
                int i = 2 * t - (t & (stride - 1));
                int j = i + stride;
                bool desc = (i & size) == 0;
                uint64_t hi_i = key_hi[i], hi_j = key_hi[j];
                uint32_t lo_i = key_lo[i], lo_j = key_lo[j];
                bool i_gt_j = key_gt<uint32_t>(hi_i, lo_i, hi_j, lo_j);
                if ((desc ? !i_gt_j : i_gt_j) && (hi_i != hi_j || lo_i != lo_j)) {
                    key_hi[i] = hi_j; key_hi[j] = hi_i;
                    key_lo[i] = lo_j; key_lo[j] = lo_i;
                }
            }
            __syncthreads();
        }
    }
    // ---- emit: `first` identifies the hit, hence the id; look the slot up again for mask ---------------
    const size_t out_base = (size_t)q * n_lists * k_max;
    for (int i = tid; i < n; i += RRF_THREADS) {
        const int first = (int)(~key_lo[i]);
        const int l = first / k_max, r = first % k_max;
        const long long id = list_ids[((size_t)l * n_q + q) * k_max + r];
        uint32_t h = rrf_hash(id, hmask);
        while (table[h].id != id) h = rrf_next(h, hmask);
        out_ids[out_base + i] = id;
        out_scores[out_base + i] = table[h].score;
        out_mask[out_base + i] = table[h].mask;
        out_first[out_base + i] = first;
    }
    for (int i = n + tid; i < n_lists * k_max; i += RRF_THREADS) {
        out_ids[out_base + i] = -1;
        out_scores[out_base + i] = -CUDART_INF;
        out_mask[out_base + i] = 0;
        out_first[out_base + i] = -1;
    }
    if (tid == 0) out_n[q] = n;
}


// The tail of the fusion stage for a batch (reference retrieval.py:485-491 the fused order, :512-516 the MMR picks, :322-333
// [:top_k] and the per-hit tags): slot j of query q takes fused position  picks[q][j]  where the query diversifies and  j
// otherwise, and the columns of that fused entry are gathered -- row id, fused score, method mask, the method whose hit
// supplies the payload and that method's own score (original_score).  One thread per (query, slot); replaces ~20 small
// tensor operations per batch.
__global__ void fuse_select_kernel(const int64_t* __restrict__ fused_ids, const double* __restrict__ fused_scores,
                                   const int32_t* __restrict__ fused_mask, const int32_t* __restrict__ fused_first,
                                   const int32_t* __restrict__ fused_n, int n_queries, int tot, const int32_t* __restrict__ picks,
                                   const int32_t* __restrict__ use_mmr, const int32_t* __restrict__ top_k,
                                   const double* __restrict__ list_scores, int k_max, int t_max, int64_t* __restrict__ out_rows,
                                   double* __restrict__ out_scores, int32_t* __restrict__ out_mask, int32_t* __restrict__ out_first_method,
                                   double* __restrict__ out_original, int32_t* __restrict__ out_n) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_queries * t_max) return;
    const int q = idx / t_max, j = idx % t_max;
    const int n_out = min(fused_n[q], top_k[q]);
    if (j == 0) out_n[q] = n_out;
    int pos = j;
    if (picks && use_mmr && use_mmr[q]) pos = max(picks[(size_t)q * t_max + j], 0);
    pos = min(pos, tot - 1);
    const size_t at = (size_t)q * tot + pos;
    const bool valid = j < n_out;
    out_rows[idx] = valid ? fused_ids[at] : -1;
    out_scores[idx] = valid ? fused_scores[at] : -CUDART_INF;
    out_mask[idx] = valid ? fused_mask[at] : 0;
    const int first = max(fused_first[at], 0);                 // list * k_max + rank0 of the hit that supplies the payload
    const int list = first / k_max;
    out_first_method[idx] = list;
    out_original[idx] = list_scores[((size_t)list * n_queries + q) * k_max + (first - list * k_max)];
}

}  // namespace b200rag

using namespace b200rag;

extern "C" {

size_t b200rag_rrf_fuse_workspace_bytes(int32_t, int32_t, int32_t) { return 256; }

int b200rag_rrf_fuse(const int64_t* list_ids, const int32_t* list_len, int32_t n_lists, int32_t n_queries,
                     int32_t k_max, const double* weights, int32_t rrf_k,
                     int64_t* out_ids, double* out_scores, int32_t* out_mask, int32_t* out_first, int32_t* out_n,
                     void* workspace, size_t workspace_bytes, void* stream) {
    (void)workspace; (void)workspace_bytes;
    B200_REQUIRE(list_ids && list_len && weights && out_ids && out_scores && out_mask && out_first && out_n,
                 "rrf_fuse: null pointer");
    B200_REQUIRE(n_lists >= 1 && n_lists <= 8 && n_queries >= 0 && k_max >= 1 && rrf_k >= 0, "rrf_fuse: bad sizes");
    if (n_queries == 0) return B200RAG_OK;
    const int total = n_lists * k_max;
    int table_size = total + total / 2 + 7;   // load factor <= 2/3
    int sort_cap = next_pow2_int(total);
    size_t smem = (size_t)table_size * sizeof(RrfSlot) + (size_t)sort_cap * 12 + 64;
    if (smem > 225 * 1024) {
        set_error("rrf_fuse: n_lists*k_max=%d needs %zu bytes of shared memory", total, smem);
        return B200RAG_E_UNSUPPORTED;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    B200_CUDA_CHECK(cudaFuncSetAttribute(rrf_fuse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rrf_fuse_kernel<<<n_queries, RRF_THREADS, smem, st>>>(list_ids, list_len, n_lists, n_queries, k_max, weights, rrf_k,
                                                         table_size, sort_cap, out_ids, out_scores, out_mask, out_first, out_n); count_launch();
    B200_CUDA_CHECK(cudaGetLastError());
    return B200RAG_OK;
}

int b200rag_fuse_select(const int64_t* fused_ids, const double* fused_scores, const int32_t* fused_mask, const int32_t* fused_first,
                        const int32_t* fused_n, int32_t n_queries, int32_t tot, const int32_t* picks, const int32_t* use_mmr,
                        const int32_t* top_k, const double* list_scores, int32_t n_lists, int32_t k_max, int32_t t_max,
                        int64_t* out_rows, double* out_scores, int32_t* out_mask, int32_t* out_first_method, double* out_original,
                        int32_t* out_n, void* stream) {
    B200_REQUIRE(fused_ids && fused_scores && fused_mask && fused_first && fused_n && top_k && list_scores && out_rows &&
                 out_scores && out_mask && out_first_method && out_original && out_n, "fuse_select: null pointer");
    B200_REQUIRE(n_queries >= 0 && n_lists >= 1 && k_max >= 1 && tot == n_lists * k_max && t_max >= 1, "fuse_select: bad sizes");
    if (n_queries == 0) return B200RAG_OK;
    const long long total = (long long)n_queries * t_max;
    fuse_select_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        fused_ids, fused_scores, fused_mask, fused_first, fused_n, n_queries, tot, picks, use_mmr, top_k, list_scores, k_max, t_max,
        out_rows, out_scores, out_mask, out_first_method, out_original, out_n);
    count_launch();
    B200_CUDA_CHECK(cudaGetLastError());
    return B200RAG_OK;
}

}  // extern "C"
