// mmr_inv.cu -- greedy MMR on token-set Jaccard, third kernel generation: heavy-token bitsets + light-token inverted lists  (K6).
//
// Same arithmetic and tie rule as mmr_select.cu (reference src/advanced_rag/retrieval.py:493-516).  What changes is how
// |tokens(c) & tokens(pick)| is found for every candidate c after a pick.  The first two generations probe, for every alive
// candidate, each of its ~90 tokens against a bitset of the picked document: n x len probes per pick (10^5 at config 4).
// A first inverted-list version (candidate lists per token, counting sort through global memory) cut the per-pick work to
// sum_c |tokens(c) & tokens(pick)| ~ 28K counter increments but paid 0.8 ms per query to build the lists: a handful of very
// common tokens sit in almost every candidate, so both the build's atomics and the per-pick increments pile up on them
// (measured: profiles/r2_mmr.md).  This version splits the query's vocabulary in two:
//
//   heavy   tokens seen in at least two of the first 32 candidates (the sample runs in shared memory; in a Zipf corpus that is
//           the ~300 tokens held by more than a few percent of the candidates), at most 384 of them.  Every candidate keeps a
//           384-bit set of its heavy tokens; |heavy(c) & heavy(pick)| is twelve AND + POPC per candidate per pick, no atomics.
//   light   every other token.  Their (token, candidate) incidences -- ~70 per candidate -- are bucketed by a hash of the token
//           into 4096 lists (counting sort: the counters live in shared memory and a light token is by construction rare, so
//           nothing piles up; the entries  token << 10 | candidate  go to a global workspace that stays in L2).  After a pick,
//           a warp per light token of the picked document reads that token's bucket (one coalesced line, ~17 entries) and bumps
//           the counter of every candidate whose entry carries the same token: sum_c |light(c) & light(pick)| ~ 300 increments.
//   Which tokens are called heavy never changes a result -- the two parts always add up to the exact intersection -- it only
//   moves work between the two mechanisms, so the sample needs no guarantee.
//   The running max similarity is kept as an integer fraction: inter/union > num/den  <=>  inter*den > num*union, exact in
//   64-bit integers and equivalent to comparing the correctly rounded fp64 quotients (distinct fractions with denominators
//   < 2^17 differ by more than 2^-34); the fp64 division -- ~50 instructions -- runs only when a candidate's maximum changes.
//
// One 1024-thread CTA per query, thread = candidate; ~100 KB of shared memory at a 100K-token vocabulary, so two CTAs share an
// SM.  Queries the scheme cannot hold (more incidences than the workspace has room for) are marked and served by the bitset
// kernels of mmr_select.cu, which are launched behind this one and leave at once for every other query.
#include "common.cuh"

namespace b200rag {

constexpr int MI_THREADS = 1024;
constexpr int MI_WARPS = MI_THREADS / 32;
constexpr unsigned MI_FULL = 0xffffffffu;
constexpr int MI_H_WORDS = 12;              // heavy-token bitset words per candidate
constexpr int MI_H = MI_H_WORDS * 32;       // heavy tokens per query (384)
constexpr int MI_SAMPLE = 32;               // candidates sampled for the heavy set (one per warp)
constexpr int MI_NB_LOG = 12;
constexpr int MI_NB = 1 << MI_NB_LOG;       // hash buckets of the light-token incidences
constexpr int MI_CAND_BITS = 10;            // entry = token << 10 | candidate  (token ids < 2^22)
constexpr int MI_TOK_UNROLL = 4;            // tokens a lane (build) / a warp (pick) has in flight

__device__ __forceinline__ uint32_t mi_bucket(int t) { return ((uint32_t)t * 2654435761u) >> (32 - MI_NB_LOG); }

// exclusive block scan of one int per thread (1024 threads); returns the prefix, *total = sum.  Two barriers.
__device__ __forceinline__ int mi_block_scan(int v, int* s_wsum, int* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(MI_FULL, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) s_wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        const int w = s_wsum[lane];
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(MI_FULL, wi, o);
            if (lane >= o) wi += u;
        }
        s_wsum[lane] = wi - w;
        if (lane == 31) s_wsum[32] = wi;
    }
    __syncthreads();
    *total = s_wsum[32];
    return incl - v + s_wsum[warp];
}

__global__ void __launch_bounds__(MI_THREADS, 2)
mmr_select_inv_kernel(const int32_t* __restrict__ cand_doc, const double* __restrict__ cand_rel, const int32_t* __restrict__ cand_n,
                      int n_max, const int64_t* __restrict__ doc_tok_ptr, const int32_t* __restrict__ doc_tok_ids, int vocab_words,
                      const double* __restrict__ lambda, const int32_t* __restrict__ k_sel, int k_max,
                      int32_t* __restrict__ out_pick, int32_t* __restrict__ out_n, uint32_t* __restrict__ ws_ent, int t_cap) {
    extern __shared__ __align__(16) char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = blockIdx.x;
    long long* s_tb = reinterpret_cast<long long*>(smem);                         // [MI_THREADS] first token of candidate c
    int* s_len = reinterpret_cast<int*>(s_tb + MI_THREADS);                       // [MI_THREADS] tokens of candidate c
    uint32_t* inter = reinterpret_cast<uint32_t*>(s_len + MI_THREADS);            // [MI_THREADS] light-token intersection counters
    uint32_t* bstart = inter + MI_THREADS;                                        // [MI_NB] bucket counts -> starts -> ends
    uint32_t* hb = bstart + MI_NB;                                                // [MI_THREADS][MI_H_WORDS] heavy sets
    uint32_t* seen1 = hb;                                                         // [vocab_words] sample: token seen (aliases hb, dead before hb is filled)
    uint32_t* heavy = hb + MI_H_WORDS * MI_THREADS;                               // [vocab_words] sample: token seen twice = heavy
    uint16_t* hpre = reinterpret_cast<uint16_t*>(heavy + vocab_words);            // [vocab_words] heavy rank of a word's first token
    __shared__ int s_wsum[33];
    __shared__ unsigned long long s_key[MI_WARPS];
    __shared__ int s_best_idx[MI_WARPS];
    __shared__ int s_done;

    uint32_t* ent = ws_ent + (size_t)q * t_cap;

    const int n = min(min(cand_n[q], n_max), MI_THREADS);
    const int k = min(min(k_sel[q], k_max), n);
    if (k <= 0) {                                             // nothing to pick (k_sel = 0: the caller skips this query)
        if (tid == 0) out_n[q] = 0;
        for (int i = tid; i < k_max; i += MI_THREADS) out_pick[(size_t)q * k_max + i] = -1;
        return;
    }
    const double lam = lambda[q];
    const double one_minus = __dsub_rn(1.0, lam);
    const int32_t* docs = cand_doc + (size_t)q * n_max;

    // ---- this thread's candidate
    const int c = tid;
    double rel = 0.0;
    int len = 0;
    bool alive = c < n;
    {
        long long tbeg = 0;
        if (alive) {
            rel = cand_rel[(size_t)q * n_max + c];
            tbeg = doc_tok_ptr[docs[c]];
            len = (int)(doc_tok_ptr[docs[c] + 1] - tbeg);
        }
        s_tb[tid] = tbeg;
        s_len[tid] = len;
    }
    for (int i = tid; i < vocab_words; i += MI_THREADS) { seen1[i] = 0u; heavy[i] = 0u; }
    for (int i = tid; i < MI_NB; i += MI_THREADS) bstart[i] = 0u;
    inter[tid] = 0u;
    if (tid == 0) s_done = 0;
    int total_tok = 0;
    mi_block_scan(len, s_wsum, &total_tok);                   // (its barriers also publish the zeroed arrays)
    if (total_tok > t_cap) {                                  // not representable here: mmr_select.cu's kernels take the query
        if (tid == 0) out_n[q] = -2;
        return;
    }
    // ---- build 1: the heavy set = tokens in at least two of the first MI_SAMPLE candidates (one candidate per warp)
    if (warp < min(n, MI_SAMPLE)) {
        const long long tb = s_tb[warp];
        const int lc = s_len[warp];
        for (int i = lane; i < lc; i += 32) {
            const int t = __ldg(doc_tok_ids + tb + i);
            const uint32_t bit = 1u << (t & 31);
            if (atomicOr(&seen1[t >> 5], bit) & bit) atomicOr(&heavy[t >> 5], bit);
        }
    }
    __syncthreads();
    {   // heavy rank of a token = prefix popcount of the heavy bitmap; only the first MI_H stay heavy (the bits of the others
        // are cleared, so "heavy" is a single bit test from here on)
        const int per = (vocab_words + MI_THREADS - 1) / MI_THREADS;
        const int w0 = tid * per, w1 = min(vocab_words, w0 + per);
        int sum = 0;
        for (int w = w0; w < w1; ++w) sum += __popc(heavy[w]);
        int n_heavy = 0;
        int run = mi_block_scan(sum, s_wsum, &n_heavy);
        for (int w = w0; w < w1; ++w) {
            uint32_t hw = heavy[w];
            const int keep = max(0, MI_H - run);
            hpre[w] = (uint16_t)min(run, MI_H);
            run += __popc(hw);
            if (__popc(hw) > keep) {
                while (__popc(hw) > keep) hw &= ~(0x80000000u >> __clz(hw));
                heavy[w] = hw;
            }
        }
    }
    __syncthreads();                                          // (seen1 is dead: hb takes its place)
    for (int i = tid; i < MI_H_WORDS * MI_THREADS; i += MI_THREADS) hb[i] = 0u;
    __syncthreads();
    // ---- build 2: heavy sets + bucket counts of the light incidences (warp per candidate, MI_TOK_UNROLL loads in flight)
    for (int cc = warp; cc < n; cc += MI_WARPS) {
        const long long tb = s_tb[cc];
        const int lc = s_len[cc];
        for (int i0 = 0; i0 < lc; i0 += 32 * MI_TOK_UNROLL) {
            int t[MI_TOK_UNROLL];
#pragma unroll
            for (int u = 0; u < MI_TOK_UNROLL; ++u) {
                const int i = i0 + u * 32 + lane;
                t[u] = i < lc ? __ldg(doc_tok_ids + tb + i) : -1;
            }
#pragma unroll
            for (int u = 0; u < MI_TOK_UNROLL; ++u) {
                if (t[u] < 0) continue;
                const uint32_t hw = heavy[t[u] >> 5], bit = 1u << (t[u] & 31);
                if (hw & bit) {
                    const int r = (int)hpre[t[u] >> 5] + __popc(hw & (bit - 1u));
                    atomicOr(&hb[cc * MI_H_WORDS + (r >> 5)], 1u << (r & 31));
                } else {
                    atomicAdd(&bstart[mi_bucket(t[u])], 1u);
                }
            }
        }
    }
    __syncthreads();
    {   // counts -> exclusive starts
        constexpr int per = MI_NB / MI_THREADS;
        uint32_t cnt[per];
        int sum = 0;
#pragma unroll
        for (int j = 0; j < per; ++j) { cnt[j] = bstart[tid * per + j]; sum += (int)cnt[j]; }
        int tot = 0;
        int run = mi_block_scan(sum, s_wsum, &tot);
#pragma unroll
        for (int j = 0; j < per; ++j) { bstart[tid * per + j] = (uint32_t)run; run += (int)cnt[j]; }
    }
    __syncthreads();
    // ---- build 3: scatter the light incidences; a bucket's start advances to its END, so afterwards
    //      bucket(b) = [b ? bstart[b - 1] : 0, bstart[b])
    for (int cc = warp; cc < n; cc += MI_WARPS) {
        const long long tb = s_tb[cc];
        const int lc = s_len[cc];
        for (int i0 = 0; i0 < lc; i0 += 32 * MI_TOK_UNROLL) {
            int t[MI_TOK_UNROLL];
#pragma unroll
            for (int u = 0; u < MI_TOK_UNROLL; ++u) {
                const int i = i0 + u * 32 + lane;
                t[u] = i < lc ? __ldg(doc_tok_ids + tb + i) : -1;
            }
#pragma unroll
            for (int u = 0; u < MI_TOK_UNROLL; ++u) {
                if (t[u] < 0) continue;
                if (!((heavy[t[u] >> 5] >> (t[u] & 31)) & 1u)) {
                    const uint32_t pos = atomicAdd(&bstart[mi_bucket(t[u])], 1u);
                    ent[pos] = ((uint32_t)t[u] << MI_CAND_BITS) | (uint32_t)cc;
                }
            }
        }
    }
    __syncthreads();

    // ---- picks
    int max_num = 0, max_den = 1;                              // running max Jaccard of this candidate as a fraction
    double max_sim = 0.0;
    for (int step = 0; step < k; ++step) {
        // 1. argmax with "earliest wins".  The fp64 score becomes an order-preserving 64-bit key (-0.0 is folded into +0.0 first,
        //    as the comparison treats them equal), reduced with three redux.sync per level: max of the high words, max of the
        //    low words among the lanes that hold it, min index among the lanes that hold both.  Key 0 = not a contender (dead,
        //    or a score that does not beat the reference's -1e9 start value).  Every warp repeats the second level on the 32
        //    warp results, so one barrier publishes the pick to the whole CTA.
        unsigned long long key = 0ull;
        if (alive) {
            const double sc = step == 0 ? rel : __dsub_rn(__dmul_rn(lam, rel), __dmul_rn(one_minus, max_sim));
            if (sc > -1e9) {
                const long long b = __double_as_longlong(__dadd_rn(sc, 0.0));
                key = (unsigned long long)b ^ (b < 0 ? 0xffffffffffffffffull : 0x8000000000000000ull);
            }
        }
        {
            const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
            const unsigned m_hi = __reduce_max_sync(MI_FULL, hi);
            const unsigned m_lo = __reduce_max_sync(MI_FULL, hi == m_hi ? lo : 0u);
            const unsigned m_i = __reduce_min_sync(MI_FULL, (hi == m_hi && lo == m_lo) ? (unsigned)c : 0x7fffffffu);
            if (lane == 0) { s_key[warp] = ((unsigned long long)m_hi << 32) | m_lo; s_best_idx[warp] = (int)m_i; }
        }
        __syncthreads();
        int pick;
        {
            const unsigned long long wk = s_key[lane];
            const unsigned hi = (unsigned)(wk >> 32), lo = (unsigned)wk;
            const unsigned m_hi = __reduce_max_sync(MI_FULL, hi);
            const unsigned m_lo = __reduce_max_sync(MI_FULL, hi == m_hi ? lo : 0u);
            const unsigned m_i = __reduce_min_sync(MI_FULL, (hi == m_hi && lo == m_lo) ? (unsigned)s_best_idx[lane] : 0x7fffffffu);
            pick = (m_hi | m_lo) ? (int)m_i : 0x7fffffff;
        }
        if (pick == 0x7fffffff) break;                        // nothing beat -1e9 (the reference would fail here too)
        if (tid == 0) {
            out_pick[(size_t)q * k_max + step] = pick;
            s_done = step + 1;
        }
        if (step + 1 == k) break;
        if (c == pick) alive = false;
        // 2. light part: a warp per light token of the picked document reads that token's bucket
        const long long tb_p = s_tb[pick];
        const int len_p = s_len[pick];
        for (int j0 = 0; j0 < len_p; j0 += MI_WARPS * MI_TOK_UNROLL) {
            int t[MI_TOK_UNROLL];
            uint32_t e0[MI_TOK_UNROLL], e1[MI_TOK_UNROLL], v[MI_TOK_UNROLL];
#pragma unroll
            for (int u = 0; u < MI_TOK_UNROLL; ++u) {
                const int j = j0 + u * MI_WARPS + warp;
                t[u] = j < len_p ? __ldg(doc_tok_ids + tb_p + j) : -1;          // (one address per warp: a broadcast load)
            }
#pragma unroll
            for (int u = 0; u < MI_TOK_UNROLL; ++u) {
                e0[u] = e1[u] = 0u;
                if (t[u] >= 0 && !((heavy[t[u] >> 5] >> (t[u] & 31)) & 1u)) {
                    const uint32_t b = mi_bucket(t[u]);
                    e0[u] = b ? bstart[b - 1] : 0u;
                    e1[u] = bstart[b];
                }
                // (ld.cg: the entries were written by this CTA a moment ago; read them where they were written, in L2)
                v[u] = e0[u] + lane < e1[u] ? __ldcg(ent + e0[u] + lane) : 0xffffffffu;
            }
#pragma unroll
            for (int u = 0; u < MI_TOK_UNROLL; ++u) {
                if (v[u] != 0xffffffffu && (int)(v[u] >> MI_CAND_BITS) == t[u]) atomicAdd(&inter[v[u] & (MI_THREADS - 1)], 1u);
                for (uint32_t e = e0[u] + 32 + lane; e < e1[u]; e += 32) {       // (buckets longer than a warp: rare)
                    const uint32_t x = __ldcg(ent + e);
                    if ((int)(x >> MI_CAND_BITS) == t[u]) atomicAdd(&inter[x & (MI_THREADS - 1)], 1u);
                }
            }
        }
        // 3. heavy part: twelve AND + POPC against the pick's heavy set (a broadcast read per word)
        int in_heavy = 0;
        {
            const uint4* mine = reinterpret_cast<const uint4*>(hb + c * MI_H_WORDS);
            const uint4* theirs = reinterpret_cast<const uint4*>(hb + pick * MI_H_WORDS);
#pragma unroll
            for (int w = 0; w < MI_H_WORDS / 4; ++w) {         // (48-byte rows: 128-bit loads of 8 consecutive rows hit 32 distinct banks)
                const uint4 a = mine[w], b = theirs[w];
                in_heavy += __popc(a.x & b.x) + __popc(a.y & b.y) + __popc(a.z & b.z) + __popc(a.w & b.w);
            }
        }
        __syncthreads();
        // 4. every candidate folds its intersection into its running maximum (integer fractions; divide only on change)
        {
            const int in = in_heavy + (int)inter[tid];
            inter[tid] = 0u;
            if (alive && in > 0) {
                const int uni = len + len_p - in;             // > 0 because in > 0
                if ((long long)in * max_den > (long long)max_num * uni) {
                    max_num = in;
                    max_den = uni;
                    max_sim = __ddiv_rn((double)in, (double)uni);
                }
            }
        }
        // (the next step's first barrier separates the counter reset from the next walk)
    }
    __syncthreads();
    const int done = s_done;
    if (tid == 0) out_n[q] = done;
    for (int i = done + tid; i < k_max; i += MI_THREADS) out_pick[(size_t)q * k_max + i] = -1;
}

size_t mmr_inv_smem_bytes(int vocab_words) {
    return (size_t)MI_THREADS * (8 + 4 + 4) + (size_t)MI_NB * 4 + (size_t)MI_H_WORDS * MI_THREADS * 4 + (size_t)vocab_words * 4 +
           (size_t)(vocab_words + (vocab_words & 1)) * 2 + 64;
}

// the sample's first bitmap aliases the heavy sets, and token ids must fit the entries' 22 bits
bool mmr_inv_vocab_ok(int vocab_words) { return vocab_words <= MI_H_WORDS * MI_THREADS && vocab_words <= (1 << (32 - MI_CAND_BITS - 5)); }

// Per query IN FLIGHT: the light (token, candidate) incidences, one u32 each, t_cap of them.  Large batches are launched
// MI_WAVE queries at a time over the same region (the launches are ordered on the stream), so the workspace stops growing
// at MI_WAVE queries (1 GB at 1000 candidates) instead of 4 GB for a 4096-query batch.
constexpr int MI_WAVE = 1024;
size_t mmr_inv_workspace_bytes(int n_queries, int t_cap) {
    return align_up((size_t)(n_queries < MI_WAVE ? n_queries : MI_WAVE) * t_cap * 4, 256) + 256;
}

int launch_mmr_inv(const int32_t* cand_doc, const double* cand_rel, const int32_t* cand_n, int n_queries, int n_max,
                   const int64_t* doc_tok_ptr, const int32_t* doc_tok_ids, int vocab_words, const double* lambda, const int32_t* k_sel,
                   int k_max, int32_t* out_pick, int32_t* out_n, void* workspace, int t_cap, cudaStream_t st) {
    const size_t smem = mmr_inv_smem_bytes(vocab_words);
    B200_CUDA_CHECK(cudaFuncSetAttribute(mmr_select_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int q0 = 0; q0 < n_queries; q0 += MI_WAVE) {
        const int nq = n_queries - q0 < MI_WAVE ? n_queries - q0 : MI_WAVE;
        mmr_select_inv_kernel<<<nq, MI_THREADS, smem, st>>>(cand_doc + (size_t)q0 * n_max, cand_rel + (size_t)q0 * n_max, cand_n + q0, n_max,
                                                              doc_tok_ptr, doc_tok_ids, vocab_words, lambda + q0, k_sel + q0, k_max,
                                                              out_pick + (size_t)q0 * k_max, out_n + q0, static_cast<uint32_t*>(workspace), t_cap);
        count_launch();
        B200_CUDA_CHECK(cudaGetLastError());
    }
    return B200RAG_OK;
}

}  // namespace b200rag
