// dense_exact.cu -- CUDA-core exact flat scan in the canonical fp64 arithmetic, plus the k-way merge kernel.
//
// dense_exact_kernel is the ground-truth GPU path: one thread owns one corpus row and scores it against QT
// queries with 8 interleaved fp64 lanes (bit-identical to oracle/exact_scan.c), the block keeps a streaming
// top-k per query in shared memory (select.cuh) and emits one sorted partial list per (query, chunk).  It is
// (a) the path B200RAG_DENSE_EXACT runs, (b) the fallback for queries the tensor-core path flags, and (c) the
// reference the tensor-core kernel is tested against on the GPU.  It is HBM/L2-latency bound and not the
// throughput path.
//
// merge_topk_kernel reduces per-query candidate lists (f64 score, i64 id) to the k best, sorted
// (score desc, id asc); it finishes the exact scan, the sparse scan and the multi-GPU all-gather.
#include "common.cuh"
#include "select.cuh"

namespace b200rag {

constexpr int EX_THREADS = 256;
constexpr int EX_QT = 4;          // queries scored per thread per row

template <int DTYPE>
__global__ void __launch_bounds__(EX_THREADS)
dense_exact_kernel(const uint16_t* __restrict__ corpus, int64_t n_rows, int dim,
                   const uint16_t* __restrict__ queries, int n_q, const int32_t* __restrict__ q_list, int k, int cap,
                   int64_t rows_per_chunk, int n_chunks, int64_t id_offset,
                   double* __restrict__ part_scores, int64_t* __restrict__ part_ids,
                   const int32_t* __restrict__ n_active, int slot_base, const uint32_t* __restrict__ row_mask) {
    extern __shared__ __align__(16) char smem[];
    const int tid = threadIdx.x;
    const int chunk = blockIdx.x;
    const int q0 = blockIdx.y * EX_QT;
    // device-side gate: only launch slots slot_base + s < *n_active carry a query (lets the tensor path's fallback be
    // launched unconditionally, without a host round trip; an idle CTA leaves at once)
    if (n_active) {
        const int na = *n_active - slot_base;
        if (na < n_q) n_q = na;
    }
    if (q0 >= n_q) return;
    const int64_t r_begin = (int64_t)chunk * rows_per_chunk;
    const int64_t r_end = min(n_rows, r_begin + rows_per_chunk);

    double* qd = reinterpret_cast<double*>(smem);          // [EX_QT][dim]
    char* p = smem + (size_t)EX_QT * dim * sizeof(double);
    BlockTopK<EX_THREADS, uint32_t> tk[EX_QT];
#pragma unroll
    for (int j = 0; j < EX_QT; ++j) {
        p = tk[j].attach(p, cap, k, EX_THREADS);
        tk[j].init();
    }
    int qidx[EX_QT];
#pragma unroll
    for (int j = 0; j < EX_QT; ++j) {
        int slot = q0 + j;
        qidx[j] = slot < n_q ? (q_list ? q_list[slot] : slot) : -1;
    }
    for (int i = tid; i < EX_QT * dim; i += EX_THREADS) {
        int j = i / dim, d = i % dim;
        qd[i] = qidx[j] >= 0 ? bits_to_double<DTYPE>(queries[(size_t)qidx[j] * dim + d]) : 0.0;
    }
    __syncthreads();

    const int64_t span = r_end - r_begin;
    const int64_t iters = (span + EX_THREADS - 1) / EX_THREADS;
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t local = it * EX_THREADS + tid;
        bool valid = local < span;
        if (valid && row_mask) valid = (__ldg(row_mask + ((r_begin + local) >> 5)) >> ((r_begin + local) & 31)) & 1u;    // metadata filter
        double s[EX_QT];
        if (valid) {
            const uint4* x = reinterpret_cast<const uint4*>(corpus + (r_begin + local) * dim);
            double acc[EX_QT][8];
#pragma unroll
            for (int j = 0; j < EX_QT; ++j)
#pragma unroll
                for (int l = 0; l < 8; ++l) acc[j][l] = 0.0;
            for (int c = 0; c < dim / 8; ++c) {
                uint4 v = __ldg(x + c);
                double xv[8];
                unpack2<DTYPE>(v.x, xv[0], xv[1]);
                unpack2<DTYPE>(v.y, xv[2], xv[3]);
                unpack2<DTYPE>(v.z, xv[4], xv[5]);
                unpack2<DTYPE>(v.w, xv[6], xv[7]);
#pragma unroll
                for (int j = 0; j < EX_QT; ++j) {
                    const double* qq = qd + (size_t)j * dim + c * 8;
#pragma unroll
                    for (int l = 0; l < 8; ++l) acc[j][l] = fma(qq[l], xv[l], acc[j][l]);
                }
            }
#pragma unroll
            for (int j = 0; j < EX_QT; ++j)
                s[j] = __dadd_rn(__dadd_rn(__dadd_rn(acc[j][0], acc[j][1]), __dadd_rn(acc[j][2], acc[j][3])),
                                 __dadd_rn(__dadd_rn(acc[j][4], acc[j][5]), __dadd_rn(acc[j][6], acc[j][7])));
        }
#pragma unroll
        for (int j = 0; j < EX_QT; ++j) tk[j].offer(valid && qidx[j] >= 0, valid ? mono64(s[j]) : 0, ~(uint32_t)local);
        // one barrier pair for all EX_QT buffers
        __syncthreads();
        int cnt[EX_QT];
#pragma unroll
        for (int j = 0; j < EX_QT; ++j) cnt[j] = tk[j].st->count;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < EX_QT; ++j)
            if (cnt[j] > cap - EX_THREADS) tk[j].compact();
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < EX_QT; ++j) {
        tk[j].finalize();
        if (qidx[j] < 0) continue;
        const int n = tk[j].count();
        const uint64_t* oh = tk[j].out_hi();
        const uint32_t* ol = tk[j].out_lo();
        double* ps = part_scores + ((size_t)(q0 + j) * n_chunks + chunk) * k;
        int64_t* pi = part_ids + ((size_t)(q0 + j) * n_chunks + chunk) * k;
        for (int i = tid; i < k; i += EX_THREADS) {
            if (i < n) {
                ps[i] = unmono64(oh[i]);
                pi[i] = id_offset + r_begin + (int64_t)(~ol[i]);
            } else {
                ps[i] = -CUDART_INF;
                pi[i] = -1;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
constexpr int MG_THREADS = 256;

template <typename OutT>
__global__ void __launch_bounds__(MG_THREADS)
merge_topk_kernel(const double* __restrict__ cand_scores, const int64_t* __restrict__ cand_ids, int n_cand, int k, int cap,
                  const int32_t* __restrict__ q_list, OutT* __restrict__ out_scores, int64_t* __restrict__ out_ids,
                  int32_t* __restrict__ out_counts, const int32_t* __restrict__ n_active, int slot_base) {
    extern __shared__ __align__(16) char smem[];
    const int tid = threadIdx.x;
    const int slot = blockIdx.x;                       // candidate lists are indexed by launch slot
    if (n_active && slot_base + slot >= *n_active) return;
    const int q = q_list ? q_list[slot] : slot;         // results go to the original query row
    BlockTopK<MG_THREADS, uint64_t> tk;
    tk.attach(smem, cap, k, MG_THREADS);
    tk.init();
    __syncthreads();
    const double* cs = cand_scores + (size_t)slot * n_cand;
    const int64_t* ci = cand_ids + (size_t)slot * n_cand;
    for (int base = 0; base < n_cand; base += MG_THREADS) {
        int i = base + tid;
        bool valid = i < n_cand;
        int64_t id = valid ? ci[i] : -1;
        valid = valid && id >= 0;
        tk.offer(valid, valid ? mono64(cs[i]) : 0, ~(uint64_t)id);
        tk.settle();
    }
    __syncthreads();
    tk.finalize();
    const int n = tk.count();
    const uint64_t* oh = tk.out_hi();
    const uint64_t* ol = tk.out_lo();
    for (int i = tid; i < k; i += MG_THREADS) {
        if (i < n) {
            out_scores[(size_t)q * k + i] = (OutT)unmono64(oh[i]);
            out_ids[(size_t)q * k + i] = (int64_t)(~ol[i]);
        } else {
            out_scores[(size_t)q * k + i] = (OutT)(-CUDART_INF);
            out_ids[(size_t)q * k + i] = -1;
        }
    }
    if (out_counts && tid == 0) out_counts[q] = n;
}

// Merge straight out of the all-gather buffer: gathered i64 [n_ranks][2][n_queries][k] holds, per rank, a plane of fp64
// score bit patterns and a plane of ids -- the two output arrays of the rank's local search, which writes them directly
// into its send buffer (b200rag/distributed.py; no pack pass between the search and the collective).  Same ranking rule.
__global__ void __launch_bounds__(MG_THREADS)
merge_gathered_kernel(const int64_t* __restrict__ gathered, int n_ranks, int n_queries, int k, int cap,
                      double* __restrict__ out_scores, int64_t* __restrict__ out_ids, int32_t* __restrict__ out_counts) {
    extern __shared__ __align__(16) char smem[];
    const int tid = threadIdx.x;
    const int q = blockIdx.x;
    BlockTopK<MG_THREADS, uint64_t> tk;
    tk.attach(smem, cap, k, MG_THREADS);
    tk.init();
    __syncthreads();
    const int n_cand = n_ranks * k;
    for (int base = 0; base < n_cand; base += MG_THREADS) {
        const int i = base + tid;
        bool valid = i < n_cand;
        int64_t id = -1;
        double sc = 0.0;
        if (valid) {
            const int g = i / k, j = i - g * k;
            const int64_t* plane = gathered + (size_t)g * 2 * n_queries * k + (size_t)q * k + j;
            id = plane[(size_t)n_queries * k];
            sc = __longlong_as_double(plane[0]);
        }
        valid = valid && id >= 0;
        tk.offer(valid, valid ? mono64(sc) : 0, ~(uint64_t)id);
        tk.settle();
    }
    __syncthreads();
    tk.finalize();
    const int n = tk.count();
    const uint64_t* oh = tk.out_hi();
    const uint64_t* ol = tk.out_lo();
    for (int i = tid; i < k; i += MG_THREADS) {
        out_scores[(size_t)q * k + i] = i < n ? unmono64(oh[i]) : -CUDART_INF;
        out_ids[(size_t)q * k + i] = i < n ? (int64_t)(~ol[i]) : -1;
    }
    if (out_counts && tid == 0) out_counts[q] = min(n, k);
}

// ------------------------------------------------------------------------------------------------ host side
struct ExactPlan {
    int n_chunks;
    int64_t rows_per_chunk;
    int cap;
    size_t smem;
};

static ExactPlan plan_exact(int64_t n_rows, int dim, int n_q, int k) {
    ExactPlan pl;
    int qtiles = (n_q + EX_QT - 1) / EX_QT;
    int target = (148 * 2 + qtiles - 1) / qtiles;                     // ~2 CTAs per SM overall
    int64_t max_chunks = (n_rows + 4 * EX_THREADS - 1) / (4 * EX_THREADS);   // >= 1024 rows per chunk
    if (max_chunks < 1) max_chunks = 1;
    pl.n_chunks = (int)(target < max_chunks ? target : max_chunks);
    if (pl.n_chunks < 1) pl.n_chunks = 1;
    while ((int64_t)pl.n_chunks * k > 65536 && pl.n_chunks > 1) pl.n_chunks /= 2;
    pl.rows_per_chunk = (n_rows + pl.n_chunks - 1) / pl.n_chunks;
    if (pl.rows_per_chunk < 1) pl.rows_per_chunk = 1;
    pl.n_chunks = (int)((n_rows + pl.rows_per_chunk - 1) / pl.rows_per_chunk);
    if (pl.n_chunks < 1) pl.n_chunks = 1;
    pl.cap = BlockTopK<EX_THREADS, uint32_t>::capacity_for(k, EX_THREADS);
    pl.smem = (size_t)EX_QT * dim * sizeof(double) + EX_QT * BlockTopK<EX_THREADS, uint32_t>::smem_bytes(pl.cap) + 64;
    return pl;
}

size_t exact_workspace_bytes(int64_t n_rows, int dim, int n_q, int k) {
    ExactPlan pl = plan_exact(n_rows, dim, n_q, k);
    return align_up((size_t)n_q * pl.n_chunks * k * sizeof(double), 256) +
           align_up((size_t)n_q * pl.n_chunks * k * sizeof(int64_t), 256) + 512;
}

int launch_merge_gated(const double* cand_scores, const int64_t* cand_ids, int n_launch, const int32_t* q_list, int n_cand, int k,
                       double* out_scores_f64, float* out_scores_f32, int64_t* out_ids, int32_t* out_counts, cudaStream_t st,
                       const int32_t* n_active, int slot_base);

int launch_merge(const double* cand_scores, const int64_t* cand_ids, int n_launch, const int32_t* q_list, int n_cand, int k,
                 double* out_scores_f64, float* out_scores_f32, int64_t* out_ids, int32_t* out_counts, cudaStream_t st) {
    return launch_merge_gated(cand_scores, cand_ids, n_launch, q_list, n_cand, k, out_scores_f64, out_scores_f32, out_ids,
                              out_counts, st, nullptr, 0);
}

int launch_merge_gated(const double* cand_scores, const int64_t* cand_ids, int n_launch, const int32_t* q_list, int n_cand, int k,
                       double* out_scores_f64, float* out_scores_f32, int64_t* out_ids, int32_t* out_counts, cudaStream_t st,
                       const int32_t* n_active, int slot_base) {
    int cap = BlockTopK<MG_THREADS, uint64_t>::capacity_for(k, MG_THREADS);
    size_t smem = BlockTopK<MG_THREADS, uint64_t>::smem_bytes(cap) + 64;
    if (smem > 220 * 1024) {
        set_error("merge_topk: k=%d needs %zu bytes of shared memory", k, smem);
        return B200RAG_E_UNSUPPORTED;
    }
    if (n_launch <= 0) return B200RAG_OK;
    if (out_scores_f64) {
        B200_CUDA_CHECK(cudaFuncSetAttribute(merge_topk_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        merge_topk_kernel<double><<<n_launch, MG_THREADS, smem, st>>>(cand_scores, cand_ids, n_cand, k, cap, q_list,
                                                                     out_scores_f64, out_ids, out_counts, n_active, slot_base); count_launch();
    } else {
        B200_CUDA_CHECK(cudaFuncSetAttribute(merge_topk_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        merge_topk_kernel<float><<<n_launch, MG_THREADS, smem, st>>>(cand_scores, cand_ids, n_cand, k, cap, q_list,
                                                                    out_scores_f32, out_ids, out_counts, n_active, slot_base); count_launch();
    }
    B200_CUDA_CHECK(cudaGetLastError());
    return B200RAG_OK;
}

// Exact scan of n_launch queries.  q_list (device, may be NULL = identity) maps launch slot -> original query
// number; query vectors are read from, and results written to, the ORIGINAL query rows.
int run_exact(const void* corpus16, int64_t n_rows, int dim, int dtype, const void* queries16, int n_launch,
              const int32_t* q_list, int k, int64_t id_offset, double* out_scores, int64_t* out_ids,
              void* workspace, size_t workspace_bytes, cudaStream_t st, const int32_t* n_active, int slot_base,
              const uint32_t* row_mask) {
    if (n_launch <= 0) return B200RAG_OK;
    ExactPlan pl = plan_exact(n_rows, dim, n_launch, k);
    if (pl.smem > 220 * 1024) {
        set_error("dense_topk(exact): k=%d dim=%d needs %zu bytes of shared memory", k, dim, pl.smem);
        return B200RAG_E_UNSUPPORTED;
    }
    Workspace ws(workspace, workspace_bytes);
    double* part_scores = ws.take<double>((size_t)n_launch * pl.n_chunks * k);
    int64_t* part_ids = ws.take<int64_t>((size_t)n_launch * pl.n_chunks * k);
    if (!ws.ok()) {
        set_error("dense_topk(exact): workspace too small (%zu < %zu)", workspace_bytes, ws.off);
        return B200RAG_E_WORKSPACE;
    }
    dim3 grid(pl.n_chunks, (n_launch + EX_QT - 1) / EX_QT);
    const uint16_t* c = static_cast<const uint16_t*>(corpus16);
    const uint16_t* q = static_cast<const uint16_t*>(queries16);
    if (dtype == B200RAG_F16) {
        B200_CUDA_CHECK(cudaFuncSetAttribute(dense_exact_kernel<B200RAG_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
        dense_exact_kernel<B200RAG_F16><<<grid, EX_THREADS, pl.smem, st>>>(c, n_rows, dim, q, n_launch, q_list, k, pl.cap,
                                                                         pl.rows_per_chunk, pl.n_chunks, id_offset, part_scores, part_ids, n_active, slot_base, row_mask); count_launch();
    } else {
        B200_CUDA_CHECK(cudaFuncSetAttribute(dense_exact_kernel<B200RAG_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
        dense_exact_kernel<B200RAG_BF16><<<grid, EX_THREADS, pl.smem, st>>>(c, n_rows, dim, q, n_launch, q_list, k, pl.cap,
                                                                          pl.rows_per_chunk, pl.n_chunks, id_offset, part_scores, part_ids, n_active, slot_base, row_mask); count_launch();
    }
    B200_CUDA_CHECK(cudaGetLastError());
    return launch_merge_gated(part_scores, part_ids, n_launch, q_list, pl.n_chunks * k, k, out_scores, nullptr, out_ids, nullptr, st,
                              n_active, slot_base);
}

}  // namespace b200rag

using namespace b200rag;

extern "C" {

size_t b200rag_merge_topk_workspace_bytes(int32_t, int32_t, int32_t) { return 256; }

int b200rag_merge_topk(const double* cand_scores, const int64_t* cand_ids, int32_t n_queries, int32_t n_cand,
                       int32_t k, double* out_scores, int64_t* out_ids, void* workspace, size_t workspace_bytes,
                       void* stream) {
    (void)workspace; (void)workspace_bytes;
    B200_REQUIRE(cand_scores && cand_ids && out_scores && out_ids, "merge_topk: null pointer");
    B200_REQUIRE(n_queries >= 0 && n_cand >= 0 && k > 0, "merge_topk: bad sizes");
    return launch_merge(cand_scores, cand_ids, n_queries, nullptr, n_cand, k, out_scores, nullptr, out_ids, nullptr,
                        static_cast<cudaStream_t>(stream));
}

int b200rag_merge_gathered(const int64_t* gathered, int32_t n_ranks, int32_t n_queries, int32_t k,
                           double* out_scores, int64_t* out_ids, int32_t* out_counts, void* stream) {
    B200_REQUIRE(gathered && out_scores && out_ids, "merge_gathered: null pointer");
    B200_REQUIRE(n_ranks >= 1 && n_queries >= 0 && k > 0, "merge_gathered: bad sizes");
    if (n_queries == 0) return B200RAG_OK;
    int cap = BlockTopK<MG_THREADS, uint64_t>::capacity_for(k, MG_THREADS);
    size_t smem = BlockTopK<MG_THREADS, uint64_t>::smem_bytes(cap) + 64;
    if (smem > 220 * 1024) {
        set_error("merge_gathered: k=%d needs %zu bytes of shared memory", k, smem);
        return B200RAG_E_UNSUPPORTED;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    B200_CUDA_CHECK(cudaFuncSetAttribute(merge_gathered_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    merge_gathered_kernel<<<n_queries, MG_THREADS, smem, st>>>(gathered, n_ranks, n_queries, k, cap, out_scores, out_ids, out_counts); count_launch();
    B200_CUDA_CHECK(cudaGetLastError());
    return B200RAG_OK;
}

}  // extern "C"
