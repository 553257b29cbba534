// mmr_select.cu -- greedy Maximal Marginal Relevance on token-set Jaccard for a batch of queries  (K6).
//
// Restates HybridRetriever._mmr_diversify (reference src/advanced_rag/retrieval.py:493-516):
//   pick 1: argmax rel;  pick j>1: argmax  lambda*rel - (1-lambda)*max_{s in selected} J(c, s)
//   J(a,b) = |a & b| / (|a | b| or 1) on token sets; strict '>' => the earliest candidate (fused order) wins ties.
// fp64 with explicit round-to-nearest ops in Python's evaluation order; the running max over the selected set is
// kept incrementally (max is order independent), turning the reference's O(k^2 n) set rebuilds into O(k n)
// intersections.
//
// One CTA per query, up to 1024 threads.  Token sets are rows of a CSR (sorted unique token ids per document).  Per pick:
//   1. argmax over the alive candidates (thread-strided, warp shuffle + one cross-warp step; "earliest wins" on ties)
//   2. the picked document's tokens are raised in a shared-memory bitset over the vocabulary
//   3. every WARP intersects whole candidates against the bitset: the 32 lanes read 32 consecutive tokens of the
//      candidate's list (one coalesced 128-byte load), probe the bitset, and the counts are summed with one warp
//      reduction.  (Round 1 had one THREAD per candidate walking its list token by token: 26.7 ms for 256 queries x 1000
//      candidates x 100 picks; this form: see profiles/r1_hybrid_c4.md.)
//   4. the bitset is cleared again (only the picked document's words).
// Work = picks x candidates x tokens per candidate bitset probes; the shared-memory bank conflicts of the random probes
// (~3.5 wavefronts per 32 probes) are the bound.
#include <cstdlib>

#include "common.cuh"

namespace b200rag {

// mmr_inv.cu: the inverted-list kernel (third generation), the product path for n_max <= 1024 and vocabularies up to ~1M tokens
size_t mmr_inv_smem_bytes(int vocab_words);
bool mmr_inv_vocab_ok(int vocab_words);
size_t mmr_inv_workspace_bytes(int n_queries, int t_cap);
int launch_mmr_inv(const int32_t* cand_doc, const double* cand_rel, const int32_t* cand_n, int n_queries, int n_max,
                   const int64_t* doc_tok_ptr, const int32_t* doc_tok_ids, int vocab_words, const double* lambda, const int32_t* k_sel,
                   int k_max, int32_t* out_pick, int32_t* out_n, void* workspace, int t_cap, cudaStream_t st);

constexpr int MMR_MAX_THREADS = 1024;
constexpr unsigned FULL = 0xffffffffu;
constexpr int MMR_G = 4;      // candidates a warp keeps in flight
constexpr int MMR_R = 4;      // 32-token rounds of a candidate held in registers (fast path: documents of <= 128 unique tokens)

__device__ __forceinline__ bool mmr_better(double ob, int oi, double b, int bi) {
    return oi != 0x7fffffff && (bi == 0x7fffffff || ob > b || (ob == b && oi < bi));
}

__global__ void __launch_bounds__(MMR_MAX_THREADS)
mmr_select_kernel(const int32_t* __restrict__ cand_doc, const double* __restrict__ cand_rel, const int32_t* __restrict__ cand_n,
                  int n_max, const int64_t* __restrict__ doc_tok_ptr, const int32_t* __restrict__ doc_tok_ids, int vocab_words,
                  const double* __restrict__ lambda, const int32_t* __restrict__ k_sel, int k_max,
                  int32_t* __restrict__ out_pick, int32_t* __restrict__ out_n, int cache_cap, int n_hi,
                  uint32_t* __restrict__ bits_global, int only_marked) {
    extern __shared__ __align__(16) char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthreads = blockDim.x, nwarps = nthreads >> 5;
    const int q = blockIdx.x;
    if (only_marked && out_n[q] != -2) return;            // served by mmr_select_inv_kernel (mmr_inv.cu)
    double* rel = reinterpret_cast<double*>(smem);                 // [n_max]
    double* max_sim = rel + n_max;                                 // [n_max]
    int64_t* tok_begin = reinterpret_cast<int64_t*>(max_sim + n_max);   // [n_max] start of the candidate's token list
    int* tok_len = reinterpret_cast<int*>(tok_begin + n_max);      // [n_max]
    int* alive = tok_len + n_max;                                  // [n_max]
    // the vocabulary bitset: in shared memory when it fits; for very large vocabularies (> ~1.2M tokens) a per-query slice of
    // the caller's workspace (vocab_smem_words = 0).  This CTA is the only reader and writer of its slice, plain stores /
    // ld.cg loads around the pick loop's barriers keep it coherent.
    const int vocab_smem_words = bits_global ? 0 : vocab_words;
    uint32_t* bits = bits_global ? bits_global + (size_t)q * vocab_words : reinterpret_cast<uint32_t*>(alive + n_max);
    // token cache: the low 16 bits of the candidates' token ids, list after list, for as many leading candidates as
    // fit; hi_bnd[c][b] = number of tokens of candidate c whose id is < (b + 1) << 16 (lists are sorted), which gives
    // the high bits back.  Reading the lists from L2 on every pick (100 picks x 360 KB per query) was the bound.
    int* cache_off = reinterpret_cast<int*>(reinterpret_cast<uint32_t*>(alive + n_max) + vocab_smem_words);   // [n_max]
    uint16_t* hi_bnd = reinterpret_cast<uint16_t*>(cache_off + n_max);                  // [n_max][n_hi]
    uint16_t* cache = hi_bnd + (size_t)n_max * n_hi + ((n_max * n_hi) & 1);             // [cache_cap], 4-byte aligned
    __shared__ int s_ncached;
    __shared__ double s_best[MMR_MAX_THREADS / 32];
    __shared__ int s_best_idx[MMR_MAX_THREADS / 32];
    __shared__ int s_pick;
    __shared__ int s_done;

    const int n = min(cand_n[q], n_max);
    const int k = min(min(k_sel[q], k_max), n);
    const double lam = lambda[q];
    const double one_minus = __dsub_rn(1.0, lam);
    const int32_t* docs = cand_doc + (size_t)q * n_max;
    for (int i = tid; i < n; i += nthreads) {
        rel[i] = cand_rel[(size_t)q * n_max + i];
        max_sim[i] = 0.0;
        alive[i] = 1;
        const int64_t b0 = doc_tok_ptr[docs[i]];
        tok_begin[i] = b0;
        tok_len[i] = (int)(doc_tok_ptr[docs[i] + 1] - b0);
    }
    for (int i = tid; i < vocab_words; i += nthreads) bits[i] = 0u;
    if (tid == 0) s_done = 0;
    __syncthreads();
    // ---- token cache: offsets (warp 0: chunked scan), then one warp per candidate copies its list ---------------
    if (warp == 0) {
        const int per = (n + 31) / 32;
        int sum = 0;
        for (int c = lane * per; c < min(n, (lane + 1) * per); ++c) sum += tok_len[c];
        int incl = sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            int v = __shfl_up_sync(FULL, incl, off);
            if (lane >= off) incl += v;
        }
        int run = incl - sum;
        int first_miss = n;
        for (int c = lane * per; c < min(n, (lane + 1) * per); ++c) {
            const bool fits = cache_cap > 0 && run + tok_len[c] <= cache_cap && tok_len[c] < 65536;
            cache_off[c] = fits ? run : -1;
            if (!fits && first_miss == n) first_miss = c;
            run += tok_len[c];
        }
        first_miss = __reduce_min_sync(FULL, first_miss);
        if (lane == 0) s_ncached = first_miss;      // candidates [0, s_ncached) are cached
    }
    __syncthreads();
    const int n_cached = s_ncached;
    for (int c = warp; c < n_cached; c += nwarps) {
        const int32_t* toks = doc_tok_ids + tok_begin[c];
        const int len_c = tok_len[c];
        uint16_t* dst = cache + cache_off[c];
        int below[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = lane; i < len_c; i += 32) {
            const int t = __ldg(toks + i);
            dst[i] = (uint16_t)(t & 0xffff);
#pragma unroll
            for (int b = 0; b < 8; ++b) below[b] += (b < n_hi && (t >> 16) <= b) ? 1 : 0;
        }
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            if (b < n_hi) {
                const int cnt = __reduce_add_sync(FULL, below[b]);
                if (lane == 0) hi_bnd[(size_t)c * n_hi + b] = (uint16_t)cnt;
            }
        }
    }
    __syncthreads();

    // one bitset probe: shared memory, or (huge vocabularies) this query's slice of the workspace read around L1
    auto probe = [&](int t) -> int {
        const uint32_t w = bits_global ? __ldcg(bits + (t >> 5)) : bits[t >> 5];
        return (int)((w >> (t & 31)) & 1u);
    };
    for (int step = 0; step < k; ++step) {
        // ---- 1. argmax with "earliest wins" -------------------------------------------------------
        double best = -1e9;
        int best_i = 0x7fffffff;
        for (int c = tid; c < n; c += nthreads) {
            if (!alive[c]) continue;
            double s = step == 0 ? rel[c] : __dsub_rn(__dmul_rn(lam, rel[c]), __dmul_rn(one_minus, max_sim[c]));
            if (s > best) { best = s; best_i = c; }   // per-thread candidates come in increasing c
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            double ob = __shfl_down_sync(FULL, best, off);
            int oi = __shfl_down_sync(FULL, best_i, off);
            if (mmr_better(ob, oi, best, best_i)) { best = ob; best_i = oi; }
        }
        if (lane == 0) { s_best[warp] = best; s_best_idx[warp] = best_i; }
        __syncthreads();
        if (warp == 0) {
            double b = lane < nwarps ? s_best[lane] : -1e9;
            int bi = lane < nwarps ? s_best_idx[lane] : 0x7fffffff;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                double ob = __shfl_down_sync(FULL, b, off);
                int oi = __shfl_down_sync(FULL, bi, off);
                if (mmr_better(ob, oi, b, bi)) { b = ob; bi = oi; }
            }
            if (lane == 0) {
                s_pick = bi;
                if (bi != 0x7fffffff) {
                    out_pick[(size_t)q * k_max + step] = bi;
                    alive[bi] = 0;
                    s_done = step + 1;
                }
            }
        }
        __syncthreads();
        const int pick = s_pick;
        if (pick == 0x7fffffff) break;           // nothing beat -1e9 (the reference would fail here too)
        if (step + 1 == k) break;
        // ---- 2. raise the picked document's tokens -------------------------------------------------
        const int64_t ps = tok_begin[pick];
        const int len_p = tok_len[pick];
        for (int i = tid; i < len_p; i += nthreads) {
            int t = doc_tok_ids[ps + i];
            atomicOr(&bits[t >> 5], 1u << (t & 31));
        }
        __syncthreads();
        // ---- 3. |tokens(c) & tokens(pick)| for every alive candidate --------------------------------------------
        // A warp takes candidates in batches of 32 (c = cb + j * nwarps): for candidate j the 32 lanes probe 32 tokens at a
        // time and the count is reduced across the warp; lane j keeps the result, and after the batch every lane does ONE
        // fp64 division (the division is ~100 instructions: one per candidate on a single active lane dominated round 1).
        for (int cb = warp; cb < n; cb += nwarps * 32) {
            int my_inter = 0, my_len = -1;
            if (cb < n_cached) {
                // ---- 3a. cached candidates: tokens from shared memory
                for (int j = 0; j < 32; ++j) {
                    const int c = cb + j * nwarps;
                    if (c >= n) break;                                    // warp-uniform
                    if (c >= n_cached || !alive[c]) continue;             // (uncached tail of a mixed batch: 3b below)
                    const uint16_t* src = cache + cache_off[c];
                    const int len_c = tok_len[c];
                    const uint16_t* bnd = hi_bnd + (size_t)c * n_hi;
                    const int b0 = n_hi ? bnd[0] : 0x7fffffff;
                    int inter = 0;
#pragma unroll
                    for (int r = 0; r < MMR_R; ++r) {
                        const int i = lane + 32 * r;
                        if (i < len_c) {
                            int hi = i >= b0;
                            for (int bb = 1; bb < n_hi; ++bb) hi += i >= bnd[bb];
                            const int t = (hi << 16) | src[i];
                            inter += probe(t);
                        }
                    }
                    for (int i = lane + 32 * MMR_R; i < len_c; i += 32) {
                        int hi = i >= b0;
                        for (int bb = 1; bb < n_hi; ++bb) hi += i >= bnd[bb];
                        const int t = (hi << 16) | src[i];
                        inter += probe(t);
                    }
                    inter = __reduce_add_sync(FULL, inter);
                    if (lane == j) { my_inter = inter; my_len = len_c; }
                }
            }
            if (cb + 31 * nwarps >= n_cached) {
                // ---- 3b. candidates that did not fit in the cache: tokens from global memory, MMR_G lists in flight
                for (int j0 = 0; j0 < 32; j0 += MMR_G) {
                    if (cb + j0 * nwarps >= n) break;                     // warp-uniform
                    int tt[MMR_G][MMR_R];
                    int lens[MMR_G];
#pragma unroll
                    for (int g = 0; g < MMR_G; ++g) {
                        const int c = cb + (j0 + g) * nwarps;
                        const bool ok = c < n && c >= n_cached && alive[c];
                        lens[g] = ok ? tok_len[c] : -1;
                        const int32_t* toks = doc_tok_ids + (ok ? tok_begin[c] : 0);
#pragma unroll
                        for (int r = 0; r < MMR_R; ++r) {
                            const int i = lane + 32 * r;
                            tt[g][r] = i < lens[g] ? __ldg(toks + i) : -1;
                        }
                    }
#pragma unroll
                    for (int g = 0; g < MMR_G; ++g) {
                        if (lens[g] < 0) continue;                        // warp-uniform
                        int inter = 0;
#pragma unroll
                        for (int r = 0; r < MMR_R; ++r) {
                            const int t = tt[g][r];
                            if (t >= 0) inter += probe(t);
                        }
                        if (lens[g] > 32 * MMR_R) {                       // long documents: the rest of the list
                            const int32_t* toks = doc_tok_ids + tok_begin[cb + (j0 + g) * nwarps];
                            for (int i = lane + 32 * MMR_R; i < lens[g]; i += 32) {
                                const int t = __ldg(toks + i);
                                inter += probe(t);
                            }
                        }
                        inter = __reduce_add_sync(FULL, inter);
                        if (lane == j0 + g) { my_inter = inter; my_len = lens[g]; }
                    }
                }
            }
            if (my_len >= 0) {
                const int c = cb + lane * nwarps;
                const int uni = my_len + len_p - my_inter;
                const double jac = __ddiv_rn((double)my_inter, (double)(uni ? uni : 1));
                if (jac > max_sim[c]) max_sim[c] = jac;
            }
        }
        __syncthreads();
        // ---- 4. clear (the next raise happens two barriers later) ----------------------------------
        for (int i = tid; i < len_p; i += nthreads) bits[doc_tok_ids[ps + i] >> 5] = 0u;
    }
    __syncthreads();
    const int done = s_done;
    if (tid == 0) out_n[q] = done;
    for (int i = done + tid; i < k_max; i += nthreads) out_pick[(size_t)q * k_max + i] = -1;
}

// ----------------------------------------------------------------------------------------------- fast path (n <= 1024)
// One THREAD per candidate, the thread <-> candidate assignment fixed for the whole query, all per-candidate state
// (relevance, running max similarity, list length, high-bit boundaries) in registers.  Candidates are counting-sorted by
// token-list length so that the 32 candidates of a warp have (nearly) equal lengths; their lists are cached TRANSPOSED in
// shared memory -- token p of the warp's 32 candidates is contiguous -- so the per-step token load is one conflict-free
// 64-byte access and every lane does useful work: ~10 warp instructions per 32 probes, no cross-lane reduction, and the
// fp64 division runs on all lanes at once.  (The warp-per-candidate kernel above spends ~1.5 warp instructions per probe.)
constexpr int MMT_THREADS = 1024;
constexpr int MMT_LEN_BINS = 1024;

__global__ void __launch_bounds__(MMT_THREADS, 1)
mmr_select_sorted_kernel(const int32_t* __restrict__ cand_doc, const double* __restrict__ cand_rel, const int32_t* __restrict__ cand_n,
                         int n_max, const int64_t* __restrict__ doc_tok_ptr, const int32_t* __restrict__ doc_tok_ids, int vocab_words,
                         const double* __restrict__ lambda, const int32_t* __restrict__ k_sel, int k_max,
                         int32_t* __restrict__ out_pick, int32_t* __restrict__ out_n, int cache_cap, int n_hi, int only_marked) {
    extern __shared__ __align__(16) char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = blockIdx.x;
    if (only_marked && out_n[q] != -2) return;            // served by mmr_select_inv_kernel (mmr_inv.cu)
    uint32_t* bits = reinterpret_cast<uint32_t*>(smem);                              // [vocab_words]
    uint16_t* cache = reinterpret_cast<uint16_t*>(bits + vocab_words);               // [cache_cap]
    __shared__ int s_hist[MMT_LEN_BINS];
    __shared__ uint16_t s_owner[MMT_THREADS];
    __shared__ int s_wsum[32], s_glen[32], s_goff[33];
    __shared__ double s_best[32];
    __shared__ int s_best_idx[32];
    __shared__ int s_pick, s_done, s_pick_len;
    __shared__ long long s_pick_begin;

    const int n = min(min(cand_n[q], n_max), MMT_THREADS);
    const int k = min(min(k_sel[q], k_max), n);
    const double lam = lambda[q];
    const double one_minus = __dsub_rn(1.0, lam);
    const int32_t* docs = cand_doc + (size_t)q * n_max;

    // ---- counting sort of the candidates by list length -------------------------------------------------------
    s_hist[tid] = 0;
    s_owner[tid] = 0xffff;
    for (int i = tid; i < vocab_words; i += MMT_THREADS) bits[i] = 0u;
    if (tid == 0) s_done = 0;
    __syncthreads();
    int key = 0;
    if (tid < n) {
        const int64_t b0 = doc_tok_ptr[docs[tid]];
        const int len = (int)(doc_tok_ptr[docs[tid] + 1] - b0);
        key = len < MMT_LEN_BINS - 1 ? len : MMT_LEN_BINS - 1;
        atomicAdd(&s_hist[key], 1);
    }
    __syncthreads();
    {
        const int v = s_hist[tid];
        int incl = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int u = __shfl_up_sync(FULL, incl, off);
            if (lane >= off) incl += u;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = s_wsum[lane];
            int wi = w;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const int u = __shfl_up_sync(FULL, wi, off);
                if (lane >= off) wi += u;
            }
            s_wsum[lane] = wi - w;                       // exclusive prefix of the warp totals
        }
        __syncthreads();
        s_hist[tid] = incl - v + s_wsum[warp];           // first position of this length
    }
    __syncthreads();
    if (tid < n) s_owner[atomicAdd(&s_hist[key], 1)] = (uint16_t)tid;
    __syncthreads();

    // ---- this thread's candidate, for the rest of the kernel ----------------------------------------------------
    const int c = tid < n ? (int)s_owner[tid] : -1;      // original candidate index (fused order), -1: idle thread
    double rel = 0.0, max_sim = 0.0;
    long long tbeg = 0;
    int len = 0;
    bool alive = c >= 0;
    if (alive) {
        rel = cand_rel[(size_t)q * n_max + c];
        tbeg = doc_tok_ptr[docs[c]];
        len = (int)(doc_tok_ptr[docs[c] + 1] - tbeg);
    }
    // group (= warp) geometry: every list of the group is padded to the longest one
    int glen = len;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) glen = max(glen, __shfl_xor_sync(FULL, glen, off));
    if (lane == 0) s_glen[warp] = glen;
    __syncthreads();
    if (warp == 0) {
        const int w = s_glen[lane] * 32;
        int wi = w;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int u = __shfl_up_sync(FULL, wi, off);
            if (lane >= off) wi += u;
        }
        s_goff[lane] = wi - w;
        if (lane == 31) s_goff[32] = wi;
    }
    __syncthreads();
    const int goff = s_goff[warp];
    const bool cached = glen > 0 && goff + glen * 32 <= cache_cap && glen < 65536;
    // fill (warp-cooperative, coalesced reads): candidate j of the group -> column j of the transposed block
    int bnd0 = 0x7fffffff, bnd1 = 0x7fffffff;            // token p has high bits  (p >= bnd0) + (p >= bnd1)
    if (cached) {
        for (int j = 0; j < 32; ++j) {
            const long long tb = __shfl_sync(FULL, tbeg, j);
            const int lj = __shfl_sync(FULL, len, j);
            int below0 = 0, below1 = 0;
            for (int i = lane; i < lj; i += 32) {
                const int t = __ldg(doc_tok_ids + tb + i);
                cache[goff + i * 32 + j] = (uint16_t)(t & 0xffff);
                below0 += (t >> 16) <= 0;
                below1 += (t >> 16) <= 1;
            }
            below0 = __reduce_add_sync(FULL, below0);
            below1 = __reduce_add_sync(FULL, below1);
            if (lane == j) { if (n_hi >= 1) bnd0 = below0; if (n_hi >= 2) bnd1 = below1; }
        }
    }
    __syncthreads();

    for (int step = 0; step < k; ++step) {
        // ---- 1. argmax with "earliest wins" ---------------------------------------------------------------------
        double best = -1e9;
        int best_i = 0x7fffffff;
        if (alive) {
            const double sc = step == 0 ? rel : __dsub_rn(__dmul_rn(lam, rel), __dmul_rn(one_minus, max_sim));
            if (sc > best) { best = sc; best_i = c; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double ob = __shfl_xor_sync(FULL, best, off);
            const int oi = __shfl_xor_sync(FULL, best_i, off);
            if (mmr_better(ob, oi, best, best_i)) { best = ob; best_i = oi; }
        }
        if (lane == 0) { s_best[warp] = best; s_best_idx[warp] = best_i; }
        __syncthreads();
        if (warp == 0) {
            double b = s_best[lane];
            int bi = s_best_idx[lane];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double ob = __shfl_xor_sync(FULL, b, off);
                const int oi = __shfl_xor_sync(FULL, bi, off);
                if (mmr_better(ob, oi, b, bi)) { b = ob; bi = oi; }
            }
            if (lane == 0) {
                s_pick = bi;
                if (bi != 0x7fffffff) {
                    out_pick[(size_t)q * k_max + step] = bi;
                    s_done = step + 1;
                }
            }
        }
        __syncthreads();
        const int pick = s_pick;
        if (pick == 0x7fffffff) break;                    // nothing beat -1e9 (the reference would fail here too)
        if (step + 1 == k) break;
        if (c == pick) { alive = false; s_pick_len = len; s_pick_begin = tbeg; }
        __syncthreads();
        // ---- 2. raise the picked document's tokens --------------------------------------------------------------
        const long long ps = s_pick_begin;
        const int len_p = s_pick_len;
        for (int i = tid; i < len_p; i += MMT_THREADS) {
            const int t = doc_tok_ids[ps + i];
            atomicOr(&bits[t >> 5], 1u << (t & 31));
        }
        __syncthreads();
        // ---- 3. every thread intersects its own candidate -------------------------------------------------------
        if (alive) {
            int inter = 0;
            if (cached) {
                // the list is sorted, so its tokens with high bits 0 / 1 / 2 are three consecutive runs: one loop per run,
                // each probing the 2048-word slice of the bitset that belongs to its high bits (no per-token select)
                const uint16_t* col = cache + goff + lane;
                const int e0 = min(bnd0, len), e1 = min(bnd1, len);
                int p = 0;
#pragma unroll 4
                for (; p < e0; ++p) {
                    const uint32_t t = col[p * 32];
                    inter += (bits[t >> 5] >> (t & 31)) & 1u;
                }
                const uint32_t* bits1 = bits + 2048;
                for (; p < e1; ++p) {
                    const uint32_t t = col[p * 32];
                    inter += (bits1[t >> 5] >> (t & 31)) & 1u;
                }
                const uint32_t* bits2 = bits + 4096;
                for (; p < len; ++p) {
                    const uint32_t t = col[p * 32];
                    inter += (bits2[t >> 5] >> (t & 31)) & 1u;
                }
            } else {
                const int32_t* toks = doc_tok_ids + tbeg;
#pragma unroll 4
                for (int p = 0; p < len; ++p) {
                    const int t = __ldg(toks + p);
                    inter += (bits[t >> 5] >> (t & 31)) & 1u;
                }
            }
            const int uni = len + len_p - inter;
            const double jac = __ddiv_rn((double)inter, (double)(uni ? uni : 1));
            if (jac > max_sim) max_sim = jac;
        }
        __syncthreads();
        // ---- 4. clear (the next raise happens two barriers later) -------------------------------------------------
        for (int i = tid; i < len_p; i += MMT_THREADS) bits[doc_tok_ids[ps + i] >> 5] = 0u;
    }
    __syncthreads();
    const int done = s_done;
    if (tid == 0) out_n[q] = done;
    for (int i = done + tid; i < k_max; i += MMT_THREADS) out_pick[(size_t)q * k_max + i] = -1;
}

}  // namespace b200rag

using namespace b200rag;

extern "C" {

// Shared-memory plan of the general kernel (bytes): `fixed` = per-candidate state, `meta` = cache offsets + high-bit boundaries.
static size_t mmr_fixed_bytes(int n_max) { return (size_t)n_max * (8 + 8 + 8 + 4 + 4) + 64; }
// Does the vocabulary bitset fit into shared memory next to the per-candidate state?  If not it lives in the workspace.
static bool mmr_bits_in_smem(int n_max, int vocab_words) {
    return mmr_fixed_bytes(n_max) + (size_t)vocab_words * 4 + (size_t)n_max * 4 + 8 <= 200 * 1024;
}

// Incidences (candidate, token) the inverted-list kernel has room for per query: 256 tokens per candidate on average.
static int mmr_inv_t_cap(int n_max) { return n_max * 256; }
static bool mmr_inv_usable(int n_max, int vocab_words) {
    const int path = option(OPT_MMR_PATH, 0);             // 0 auto, 1 general bitset kernel, 2 bitset kernels, 3 inverted lists ONLY
    return n_max <= 1024 && mmr_inv_vocab_ok(vocab_words) && mmr_inv_smem_bytes(vocab_words) <= 200 * 1024 && (path == 0 || path == 3);
}

size_t b200rag_mmr_select_workspace_bytes(int32_t n_queries, int32_t n_max, int32_t vocab_size) {
    if (n_queries <= 0 || n_max <= 0 || vocab_size <= 0) return 256;
    const int vocab_words = (vocab_size + 31) / 32;
    if (mmr_inv_usable(n_max, vocab_words)) return mmr_inv_workspace_bytes(n_queries, mmr_inv_t_cap(n_max));
    if (mmr_bits_in_smem(n_max, vocab_words)) return 256;
    return align_up((size_t)n_queries * vocab_words * 4, 256) + 256;      // one bitset slice per query
}

int b200rag_mmr_select(const int32_t* cand_doc, const double* cand_rel, const int32_t* cand_n, int32_t n_queries,
                       int32_t n_max, const int64_t* doc_tok_ptr, const int32_t* doc_tok_ids, int32_t vocab_size,
                       const double* lambda, const int32_t* k_sel, int32_t k_max,
                       int32_t* out_pick, int32_t* out_n,
                       void* workspace, size_t workspace_bytes, void* stream) {
    B200_REQUIRE(cand_doc && cand_rel && cand_n && doc_tok_ptr && lambda && k_sel && out_pick && out_n, "mmr_select: null pointer");
    B200_REQUIRE(n_queries >= 0 && n_max >= 1 && vocab_size >= 1 && k_max >= 1, "mmr_select: bad sizes");
    if (n_queries == 0) return B200RAG_OK;
    int vocab_words = (vocab_size + 31) / 32;
    // tokens >= (n_hi << 16) do not exist; n_hi thresholds per candidate give the high bits of cached tokens back
    int n_hi = (vocab_size - 1) >> 16;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // Product path: the inverted-list kernel (mmr_inv.cu).  It marks the queries it cannot hold (out_n = -2); the bitset kernels
    // below are then launched for exactly those -- every other CTA of theirs leaves at once.  With a workspace that is too small
    // (callers of ABI version 1 passed 256 bytes) the bitset kernels serve every query, as before.
    int only_marked = 0;
    if (mmr_inv_usable(n_max, vocab_words) && workspace &&
        workspace_bytes >= mmr_inv_workspace_bytes(n_queries, mmr_inv_t_cap(n_max)) && ((uintptr_t)workspace & 255) == 0) {
        int rc = launch_mmr_inv(cand_doc, cand_rel, cand_n, n_queries, n_max, doc_tok_ptr, doc_tok_ids, vocab_words, lambda, k_sel,
                                k_max, out_pick, out_n, workspace, mmr_inv_t_cap(n_max), st);
        if (rc) return rc;
        only_marked = 1;
        if (option(OPT_MMR_PATH, 0) == 3) return B200RAG_OK;      // (tests: queries the kernel could not hold keep out_n = -2)
    }
    if (n_max >= 32 && n_max <= MMT_THREADS && n_hi <= 2 && option(OPT_MMR_PATH, 0) != 1) {
        // fast path: one thread per candidate, transposed token cache (static shared memory of the kernel: ~7 KB)
        const size_t limit = 227 * 1024 - 8192;
        const size_t bits_bytes = (size_t)vocab_words * 4;
        if (bits_bytes + 65536 <= limit) {
            const int cap = (int)((limit - bits_bytes) / 2);
            const size_t smem_t = bits_bytes + (size_t)cap * 2;
            B200_CUDA_CHECK(cudaFuncSetAttribute(mmr_select_sorted_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
            mmr_select_sorted_kernel<<<n_queries, MMT_THREADS, smem_t, st>>>(cand_doc, cand_rel, cand_n, n_max, doc_tok_ptr, doc_tok_ids,
                                                                              vocab_words, lambda, k_sel, k_max, out_pick, out_n, cap, n_hi, only_marked);
            count_launch();
            B200_CUDA_CHECK(cudaGetLastError());
            return B200RAG_OK;
        }
    }
    // general kernel.  The launch always covers `meta` (cache offsets are written for every candidate even when nothing is
    // cached -- round 1 launched without it when the cache was disabled and wrote past the allocation).
    const bool bits_smem = mmr_bits_in_smem(n_max, vocab_words);
    uint32_t* bits_global = nullptr;
    if (!bits_smem) {
        const size_t need = b200rag_mmr_select_workspace_bytes(n_queries, n_max, vocab_size);
        if (!workspace || workspace_bytes < need) {
            set_error("mmr_select: vocab=%d needs a %zu-byte workspace for the token bitsets (got %zu)", vocab_size, need, workspace_bytes);
            return B200RAG_E_WORKSPACE;
        }
        bits_global = static_cast<uint32_t*>(workspace);
    }
    const size_t fixed = mmr_fixed_bytes(n_max) + (bits_smem ? (size_t)vocab_words * 4 : 0);
    if (n_hi > 8) n_hi = 0;                            // token ids >= 2^19: no 16-bit token cache (lists are read through L2)
    size_t meta = (size_t)n_max * 4 + (size_t)n_max * n_hi * 2 + 8;
    const size_t limit = 227 * 1024 - 1024;            // static shared memory (reduction scratch) takes the rest
    int cache_cap = 0;
    if (n_hi > 0 || vocab_size <= 65536) {
        if (fixed + meta + 4096 <= limit) cache_cap = (int)((limit - fixed - meta) / 2);
    }
    if (cache_cap == 0) { n_hi = 0; meta = (size_t)n_max * 4 + 8; }
    const size_t smem = fixed + meta + (size_t)cache_cap * 2;
    if (smem > limit) {
        set_error("mmr_select: n_max=%d needs %zu bytes of shared memory", n_max, smem);
        return B200RAG_E_UNSUPPORTED;
    }
    B200_CUDA_CHECK(cudaFuncSetAttribute(mmr_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // one warp per candidate in the intersection phase: as many warps as there are candidates, up to 32
    const int threads = n_max >= 32 ? MMR_MAX_THREADS : (n_max >= 8 ? 256 : 128);
    mmr_select_kernel<<<n_queries, threads, smem, st>>>(cand_doc, cand_rel, cand_n, n_max, doc_tok_ptr, doc_tok_ids,
                                                           vocab_words, lambda, k_sel, k_max, out_pick, out_n, cache_cap, n_hi, bits_global, only_marked); count_launch();
    B200_CUDA_CHECK(cudaGetLastError());
    return B200RAG_OK;
}

}  // extern "C"
