// sparse_bm25.cu -- sparse inner-product (BM25-weighted) top-k over doc-range-blocked postings  (K3, round-2 generation).
//
// One 256-thread CTA owns one (query, slice of document blocks) and walks its blocks one after the other; two or three CTAs
// share an SM.  The block's per-document accumulators live in shared memory (fp32), zeroed once per CTA -- collecting a block
// puts every touched accumulator back to zero.  What bounds this kernel is not bytes (6 per posting) but the dependent chain
// inside a block: ranges -> postings -> accumulator read-modify-write per term, in term order.  Round 1 ran that chain once per
// CTA with a block barrier between terms and found the candidates by reading all 16384 accumulators back (7.9 warp instructions
// per posting, profiles/r1_sparse_ncu.md).  This generation:
//
//   staging     the postings of the NEXT block (rows u16 + weights f32 of up to 8 query terms) are copied into shared memory with
//               cp.async while the current block is processed (ranges are fetched two blocks ahead): no global-memory latency
//               on the chain, and the collect pass re-reads the rows from shared memory.
//   warp-private sub-ranges   every warp owns 1/8 of the block's documents.  One flat pass over the staged (doc-sorted) lists
//               records where each list crosses the sub-range boundaries; then each warp applies ITS part of the 8 lists in
//               ascending term id -- acc = fmaf(qv, w, acc), the canonical order, bit-identical to oracle/exact_scan.c -- with
//               __syncwarp only: eight independent chains per CTA instead of one, no block barrier between terms.
//   collect     proportional to the postings, not to the documents: once the running k-th best score `thr` is positive (after the
//               first block or two) each warp walks its postings again, takes the final score out of the accumulator and
//               zeroes it (later postings of the same document then read 0); only scores >= thr go on.  No scan over the
//               accumulators, no touched-bitmap.  While thr <= 0 (first block, queries with non-positive scores) a bitmap marks
//               the candidates instead -- a document with score exactly 0 that shares a term with the query is still a hit.
//   survivors   go through a staging list into the CTA's streaming top-k (select.cuh); the list is drained (which raises thr)
//               when it is half full, not after every block.  If it overflows, the overflowing documents keep their
//               accumulator, the list is drained and the collect pass of that block runs again.
// Blocks that do not fit this scheme (more staged postings than the buffer holds, queries with more than 8 terms, block sizes
// that are not a power of two >= 256) take the block-level path: same arithmetic, a block barrier between terms, postings read
// in place.  Slices of one query share their thresholds through a global atomicMax.  Algorithmic HBM traffic = 6 bytes per
// posting of the query's terms.  grid = (queries, slices); merge_topk_kernel reduces the slices.
#include <algorithm>

#include "sparse.cuh"

namespace b200rag {

int launch_merge(const double* cand_scores, const int64_t* cand_ids, int n_launch, const int32_t* q_list, int n_cand, int k,
                 double* out_scores_f64, float* out_scores_f32, int64_t* out_ids, int32_t* out_counts, cudaStream_t st);

constexpr int SP_THREADS = 256;
constexpr int SP_WARPS = SP_THREADS / 32;
constexpr int SP_TG = 8;            // query terms handled together
constexpr int SP_STAGE = 1024;      // survivors staged per collect pass
constexpr int SP_PCAP = 2560;       // postings of one block staged in shared memory (per buffer; the rest is read in place)
constexpr int SP_RING = 3;          // range / layout slots: block b uses slot b % 3, slot 3 serves the term groups beyond the first
constexpr int SP_NSTAT = 12;        // debug counters per CTA (b200rag_debug_set_stats_buffer kind 1)
constexpr int SP_STAT_CTAS = 1024;
constexpr int SP_MAX_SLICES = 32;

enum { SPS_TOTAL = 0, SPS_ACC, SPS_COLLECT, SPS_DRAIN, SPS_BLOCKS, SPS_STAGED, SPS_RESCANS, SPS_BITMAP_BLOCKS, SPS_POSTINGS,
       SPS_WAIT, SPS_UNSTAGED };

__device__ __forceinline__ void sp_cp_async4(void* smem_dst, const void* gsrc, unsigned src_bytes) {
    // 4-byte asynchronous global -> shared copy; src_bytes < 4 zero-fills the rest (nothing beyond src_bytes is read)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc),
                 "r"(src_bytes)
                 : "memory");
}

__global__ void __launch_bounds__(SP_THREADS, 2) sparse_query_kernel(const SparseParams p) {
    extern __shared__ __align__(16) char smem[];
    // per ring slot: posting ranges of a term group inside one block and where their staged copies live
    __shared__ long long s_beg[SP_RING + 1][SP_TG];
    __shared__ int s_len[SP_RING + 1][SP_TG];
    __shared__ int s_slen[SP_RING + 1][SP_TG];      // leading postings of the list that are staged
    __shared__ int s_woff[SP_RING + 1][SP_TG];      // staged weights start here (staging buffer index)
    __shared__ int s_doff[SP_RING + 1][SP_TG];      // staged rows start here (u16 index; even base + the list's odd/even shift)
    __shared__ int s_off[SP_RING + 1][SP_TG + 1];   // exclusive prefix of s_len: the flat index space of the collect pass
    __shared__ float s_qv[SP_RING + 1][SP_TG];
    __shared__ int s_fits[SP_RING + 1];             // every posting of the slot's lists is staged
    __shared__ int s_bnd[SP_TG][SP_WARPS + 1];      // current block: where list j crosses the warps' document sub-ranges
    __shared__ int s_nstage;
    __shared__ unsigned int s_gthr;
    __shared__ unsigned long long s_stat[SP_NSTAT];
    __shared__ long long s_last;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = blockIdx.x, slice = blockIdx.y;
    {   // queries of up to p.mask_max_terms terms are served by sparse_mask_kernel (sparse_mask.cu); this CTA has nothing to do
        const int nq0 = (int)(p.q_ptr[q + 1] - p.q_ptr[q]);
        if (nq0 <= p.mask_max_terms) return;
    }
    const int block_docs = p.block_docs, n_words = block_docs >> 5;
    float* acc = reinterpret_cast<float*>(smem);                                     // [block_docs]
    uint32_t* touched = reinterpret_cast<uint32_t*>(acc + block_docs);                // [n_words]
    uint32_t* stage_doc = touched + n_words;                                          // [SP_STAGE] row inside the block
    float* stage_sc = reinterpret_cast<float*>(stage_doc + SP_STAGE);                 // [SP_STAGE]
    float* pw = stage_sc + SP_STAGE;                                                  // [2][SP_PCAP] staged weights
    uint16_t* pd = reinterpret_cast<uint16_t*>(pw + 2 * SP_PCAP);                     // [2][SP_PCAP + 32] staged rows
    constexpr int PD_STRIDE = SP_PCAP + 32;
    char* tkmem = reinterpret_cast<char*>(pd + 2 * PD_STRIDE);
    tkmem = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(tkmem) + 15) & ~uintptr_t(15));
    using TopK = BlockTopK<SP_THREADS, uint32_t>;
    TopK tk;
    tk.attach(tkmem, p.cap, p.k, SP_THREADS, /*start_digit=*/TopK::NLO + 3);
    tk.init();
    for (int i = tid; i < block_docs; i += SP_THREADS) acc[i] = 0.0f;
    for (int i = tid; i < n_words; i += SP_THREADS) touched[i] = 0u;
    const bool stats = p.stats != nullptr;
    if (tid == 0) {
        s_nstage = 0;
        s_gthr = 0u;
        if (stats) {
            for (int i = 0; i < SP_NSTAT; ++i) s_stat[i] = 0;
            s_last = clock64();
        }
    }
#define SP_MARK(i)                                                       \
    do {                                                                 \
        if (stats && tid == 0) {                                         \
            const long long now_ = clock64();                            \
            s_stat[i] += (unsigned long long)(now_ - s_last);            \
            s_last = now_;                                               \
        }                                                                \
    } while (0)

    const int64_t qs = p.q_ptr[q];
    const int nq = (int)(p.q_ptr[q + 1] - qs);
    const int n_groups = (nq + SP_TG - 1) / SP_TG;
    const bool multi = n_groups > 1;
    const bool staging = !(p.flags & 2);
    const int b0 = (int)((int64_t)slice * p.n_blocks / p.n_slices), b1 = (int)((int64_t)(slice + 1) * p.n_blocks / p.n_slices);
    const size_t row_stride = (size_t)p.n_terms + 1;
    // postings in the whole index = the end pointer of the last block (bounds the 4-byte row copies at the very end)
    const long long nnz = p.blk_term_ptr[(size_t)(p.n_blocks - 1) * row_stride + p.n_terms];

    // ranges of term group g inside block blk: thread j < SP_TG owns term g * SP_TG + j
    int my_t = -1;
    float my_qv = 0.f;
    auto load_term = [&](int g) {
        my_t = -1;
        my_qv = 0.f;
        const int j = g * SP_TG + tid;
        if (tid < SP_TG && j < nq) {
            const int t = p.q_terms[qs + j];
            if (t >= 0 && t < p.n_terms) { my_t = t; my_qv = p.q_vals[qs + j]; }
        }
    };
    auto fetch_range = [&](int blk, long long& rb, int& rl) {
        rb = 0;
        rl = 0;
        if (my_t >= 0) {
            const int64_t* src = p.blk_term_ptr + (size_t)blk * row_stride + my_t;
            rb = src[0];
            rl = (int)(src[1] - rb);
        }
    };
    // Warp 0 publishes the ranges its lanes 0..7 hold into ring slot `slot`, together with the staging layout: the flat
    // prefix (collect pass), how much of every list fits the staging buffer and where.  Followed by a barrier.
    auto publish = [&](int slot, long long rb, int rl, bool stage_it) {
        if (warp != 0) return;
        int len = tid < SP_TG ? rl : 0;
        int incl = len, wincl;
#pragma unroll
        for (int o = 1; o < SP_TG; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const int excl = incl - len;
        const int woff = min(excl, SP_PCAP);
        const int slen = stage_it ? max(0, min(len, SP_PCAP - excl)) : 0;
        // rows are staged as aligned 4-byte words: a list that starts at an odd posting index keeps its shift
        const int sh = (int)(rb & 1);
        int dw = slen ? 2 * ((sh + slen + 1) >> 1) : 0;        // u16 slots this list occupies (even)
        wincl = dw;
#pragma unroll
        for (int o = 1; o < SP_TG; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, wincl, o);
            if (lane >= o) wincl += v;
        }
        if (tid < SP_TG) {
            s_beg[slot][tid] = rb;
            s_len[slot][tid] = len;
            s_slen[slot][tid] = slen;
            s_woff[slot][tid] = woff;
            s_doff[slot][tid] = wincl - dw + sh;
            s_off[slot][tid] = excl;
            s_qv[slot][tid] = my_qv;
            if (tid == SP_TG - 1) {
                s_off[slot][SP_TG] = incl;
                s_fits[slot] = stage_it && incl <= SP_PCAP;
            }
        }
    };
    // All threads: start the asynchronous copies of a published block's leading postings into staging buffer `buf`.
    auto issue_stage = [&](int slot, int buf) {
        float* wdst = pw + buf * SP_PCAP;
        uint32_t* ddst = reinterpret_cast<uint32_t*>(pd + buf * PD_STRIDE);
        const uint32_t* dsrc32 = reinterpret_cast<const uint32_t*>(p.post_doc);
#pragma unroll
        for (int j = 0; j < SP_TG; ++j) {
            const int slen = s_slen[slot][j];
            if (slen == 0) continue;
            const long long beg = s_beg[slot][j];
            const int woff = s_woff[slot][j];
            for (int i = tid; i < slen; i += SP_THREADS) sp_cp_async4(wdst + woff + i, p.post_w + beg + i, 4u);
            const int sh = (int)(beg & 1);
            const long long g0 = beg - sh;                                 // even posting index
            const int nw = (sh + slen + 1) >> 1;
            const int dbase = (s_doff[slot][j] - sh) >> 1;                 // word index inside the staging buffer
            for (int wi = tid; wi < nw; wi += SP_THREADS)
                sp_cp_async4(ddst + dbase + wi, dsrc32 + (g0 >> 1) + wi, g0 + 2 * wi + 1 < nnz ? 4u : 2u);
        }
    };

    // ---- ordered accumulate of the term group in `slot`; `prev_pw` = warps that took part in the previous non-empty step
    auto accumulate_group = [&](int slot, int buf, bool bitmap, int& prev_pw) {
        const float* wsrc = pw + buf * SP_PCAP;
        const uint16_t* dsrc = pd + buf * PD_STRIDE;
#pragma unroll
        for (int j = 0; j < SP_TG; ++j) {
            const int len = s_len[slot][j];
            if (len == 0) continue;                                                  // CTA-uniform
            const int pwarps = len >= SP_THREADS ? SP_WARPS : (len + 31) >> 5;
            if (prev_pw) {
                if (prev_pw > 1 || pwarps > 1) __syncthreads();
                else __syncwarp();
            }
            prev_pw = pwarps;
            if (warp < pwarps) {
                const float qv = s_qv[slot][j];
                const int slen = s_slen[slot][j];
                const float* ws = wsrc + s_woff[slot][j];
                const uint16_t* ds = dsrc + s_doff[slot][j];
                int i = tid;
                for (; i < slen; i += SP_THREADS) {                                  // staged part: shared memory only
                    const int d = ds[i];
                    acc[d] = fmaf(qv, ws[i], acc[d]);
                    if (bitmap) atomicOr(&touched[d >> 5], 1u << (d & 31));
                }
                const long long beg = s_beg[slot][j];
                for (; i < len; i += 2 * SP_THREADS) {                               // the rest of a long list: two in flight
                    const int i1 = i + SP_THREADS;
                    const int d0 = p.post_doc[beg + i];
                    const float w0 = p.post_w[beg + i];
                    int d1 = -1;
                    float w1 = 0.f;
                    if (i1 < len) { d1 = p.post_doc[beg + i1]; w1 = p.post_w[beg + i1]; }
                    acc[d0] = fmaf(qv, w0, acc[d0]);
                    if (bitmap) atomicOr(&touched[d0 >> 5], 1u << (d0 & 31));
                    if (d1 >= 0) {
                        acc[d1] = fmaf(qv, w1, acc[d1]);
                        if (bitmap) atomicOr(&touched[d1 >> 5], 1u << (d1 & 31));
                    }
                }
            }
        }
    };

    // a survivor goes to the staging list; false = the list is full (the caller keeps the document for the next pass)
    auto stage = [&](uint32_t doc, float sc) -> bool {       // doc = row inside this shard (block start + row in block)
        const int slot = atomicAdd(&s_nstage, 1);
        if (slot >= SP_STAGE) return false;
        stage_doc[slot] = doc;
        stage_sc[slot] = sc;
        return true;
    };

    // ---- exchange collect over the term group in `slot`: flat index space, every posting visited once
    auto collect_group_exch = [&](int slot, int buf, float thr_f, int64_t doc0) {
        const uint16_t* dsrc = pd + buf * PD_STRIDE;
        const int total = s_off[slot][SP_TG];
        int j = 0;
        for (int pos = tid; pos < total; pos += SP_THREADS) {
            while (pos >= s_off[slot][j + 1]) ++j;
            const int i = pos - s_off[slot][j];
            const int d = i < s_slen[slot][j] ? (int)dsrc[s_doff[slot][j] + i] : (int)p.post_doc[s_beg[slot][j] + i];
            const float sc = atomicExch(&acc[d], 0.0f);
            if (sc >= thr_f) {
                bool ok = true;
                if (p.doc_mask) {
                    const int64_t g = doc0 + d;
                    ok = (__ldg(p.doc_mask + (g >> 5)) >> (g & 31)) & 1u;
                }
                if (ok && !stage((uint32_t)(doc0 + d), sc)) acc[d] = sc;     // list full: the document waits for the next pass
            }
        }
    };
    // ---- bitmap collect: every thread owns whole 32-document words
    auto collect_bitmap = [&](float thr_f, int64_t doc0, int w_begin, int w_end, int w_first, int w_step) {
        const int64_t n_mask_words = (p.n_docs + 31) >> 5;
        for (int w = w_begin + w_first; w < w_end; w += w_step) {
            uint32_t m = touched[w];
            if (!m) continue;
            uint32_t allowed = 0xffffffffu;
            if (p.doc_mask) {                                  // block_docs % 32 == 0: a bitmap word is a word of the mask
                const int64_t gw = (doc0 >> 5) + w;
                allowed = gw < n_mask_words ? __ldg(p.doc_mask + gw) : 0u;
            }
            uint32_t keep = 0;
            while (m) {
                const int b = __ffs((int)m) - 1;
                m &= m - 1;
                const int d = w * 32 + b;
                const float sc = acc[d];
                if (!(sc < thr_f) && ((allowed >> b) & 1u)) {
                    if (stage((uint32_t)(doc0 + d), sc)) acc[d] = 0.0f;
                    else keep |= 1u << b;
                } else {
                    acc[d] = 0.0f;
                }
            }
            touched[w] = keep;
        }
    };

    // ---- prologue: ranges of the first block published and staged, ranges of the second one in flight
    load_term(0);
    long long nb = 0;
    int nl = 0;
    const bool walk = nq > 0 && b0 < b1;
    if (walk) fetch_range(b0, nb, nl);
    __syncthreads();                                           // (init of acc / touched / top-k done)
    if (walk) {
        publish(b0 % SP_RING, nb, nl, staging);
        __syncthreads();
        issue_stage(b0 % SP_RING, b0 & 1);
        if (b0 + 1 < b1) fetch_range(b0 + 1, nb, nl);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    SP_MARK(SPS_TOTAL);
    // warp-private sub-ranges need a power-of-two block of at least 32 documents per warp
    const bool fast_geom = (block_docs & (block_docs - 1)) == 0 && block_docs >= 32 * SP_WARPS;
    const int sub_shift = 31 - __clz(block_docs) - 3;             // log2(block_docs / SP_WARPS)
    const int sub_words = n_words / SP_WARPS;
    float thr_f = -CUDART_INF_F;
    for (int blk = b0; blk < b1 && nq > 0; ++blk) {
        const int64_t doc0 = (int64_t)blk * block_docs;
        const int slot = blk % SP_RING, buf = blk & 1;
        // ---- block top: publish + stage the NEXT block, exchange thresholds with the other slices, wait for THIS block's copies
        if (blk + 1 < b1) publish((blk + 1) % SP_RING, nb, nl, staging);
        if (tid == 0 && p.gthr) {
            if (tk.st->has_thr) atomicMax(p.gthr + q, (unsigned int)tk.st->thr_hi);
            unsigned int g;
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(g) : "l"(p.gthr + q) : "memory");
            s_gthr = g;
        }
        __syncthreads();
        if (blk + 1 < b1) issue_stage((blk + 1) % SP_RING, buf ^ 1);
        asm volatile("cp.async.commit_group;" ::: "memory");   // (one group per block, possibly empty)
        if (blk + 2 < b1) fetch_range(blk + 2, nb, nl);        // (in flight during this whole block)
        asm volatile("cp.async.wait_group 1;" ::: "memory");   // everything but the newest group has landed: this block's
        __syncthreads();
        SP_MARK(SPS_WAIT);
        if (s_gthr) thr_f = fmaxf(thr_f, unmono32(s_gthr));
        const bool bitmap = !(thr_f > 0.0f) || (p.flags & 1);  // CTA-uniform
        const bool fast = fast_geom && !multi && s_fits[slot] && !(p.flags & 4);
        const float* wsrc = pw + buf * SP_PCAP;
        const uint16_t* dsrc = pd + buf * PD_STRIDE;
        if (fast) {
            // ---- where does each (doc-sorted) staged list cross the warps' sub-range boundaries?  One flat pass.
            {
                const int total = s_off[slot][SP_TG];
                int j = 0;
                for (int pos = tid; pos < total; pos += SP_THREADS) {
                    while (pos >= s_off[slot][j + 1]) ++j;
                    const int i = pos - s_off[slot][j];
                    const uint16_t* ds = dsrc + s_doff[slot][j];
                    const int wcur = ds[i] >> sub_shift;
                    const int wprev = i ? (ds[i - 1] >> sub_shift) : -1;
                    for (int w = wprev + 1; w <= wcur; ++w) s_bnd[j][w] = i;
                    if (i == s_len[slot][j] - 1)
                        for (int w = wcur + 1; w <= SP_WARPS; ++w) s_bnd[j][w] = i + 1;
                }
                if (tid < SP_TG && s_len[slot][tid] == 0)
                    for (int w = 0; w <= SP_WARPS; ++w) s_bnd[tid][w] = 0;
            }
            __syncthreads();
            // ---- accumulate: this warp's documents, all lists in ascending term order, warp-level ordering only
#pragma unroll
            for (int j = 0; j < SP_TG; ++j) {
                const int lo = s_bnd[j][warp], hi = s_bnd[j][warp + 1];
                if (lo < hi) {                                 // warp-uniform
                    const float qv = s_qv[slot][j];
                    const float* ws = wsrc + s_woff[slot][j];
                    const uint16_t* ds = dsrc + s_doff[slot][j];
                    for (int i = lo + lane; i < hi; i += 32) {
                        const int d = ds[i];
                        acc[d] = fmaf(qv, ws[i], acc[d]);
                        if (bitmap) atomicOr(&touched[d >> 5], 1u << (d & 31));
                    }
                    __syncwarp();
                }
            }
            SP_MARK(SPS_ACC);
        } else {
            // ---- block-level path: all term groups in ascending term order, a block barrier between lists
            int prev_pw = 0;
            accumulate_group(slot, buf, bitmap, prev_pw);
            for (int g = 1; g < n_groups; ++g) {               // queries with more than SP_TG terms: unstaged, unpipelined
                __syncthreads();
                load_term(g);
                long long rb;
                int rl;
                fetch_range(blk, rb, rl);
                publish(SP_RING, rb, rl, false);
                __syncthreads();
                prev_pw = 0;                                   // (the barrier above already ordered the previous group)
                accumulate_group(SP_RING, buf, bitmap, prev_pw);
            }
            __syncthreads();
            SP_MARK(SPS_ACC);
        }
        if (stats && tid == 0) {
            s_stat[SPS_BLOCKS] += 1;
            s_stat[SPS_BITMAP_BLOCKS] += bitmap ? 1 : 0;
            s_stat[SPS_POSTINGS] += (unsigned long long)s_off[slot][SP_TG];
            s_stat[SPS_UNSTAGED] += fast ? 0 : 1;
        }
        // ---- collect (repeated while the staging list overflows)
        for (;;) {
            if (fast) {
                if (bitmap) {
                    collect_bitmap(thr_f, doc0, warp * sub_words, (warp + 1) * sub_words, lane, 32);
                } else {
#pragma unroll
                    for (int j = 0; j < SP_TG; ++j) {
                        const int lo = s_bnd[j][warp], hi = s_bnd[j][warp + 1];
                        if (lo < hi) {
                            const uint16_t* ds = dsrc + s_doff[slot][j];
                            for (int i = lo + lane; i < hi; i += 32) {
                                const int d = ds[i];
                                const float sc = acc[d];       // the document's final score; 0 if an earlier list took it
                                acc[d] = 0.0f;
                                if (sc >= thr_f) {
                                    bool ok = true;
                                    if (p.doc_mask) {
                                        const int64_t g = doc0 + d;
                                        ok = (__ldg(p.doc_mask + (g >> 5)) >> (g & 31)) & 1u;
                                    }
                                    if (ok && !stage((uint32_t)(doc0 + d), sc)) acc[d] = sc;   // list full: next pass
                                }
                            }
                            __syncwarp();
                        }
                    }
                }
            } else if (bitmap) {
                collect_bitmap(thr_f, doc0, 0, n_words, tid, SP_THREADS);
            } else {
                collect_group_exch(slot, buf, thr_f, doc0);
                for (int g = 1; g < n_groups; ++g) {
                    __syncthreads();
                    load_term(g);
                    long long rb;
                    int rl;
                    fetch_range(blk, rb, rl);
                    publish(SP_RING, rb, rl, false);
                    __syncthreads();
                    collect_group_exch(SP_RING, buf, thr_f, doc0);
                }
            }
            __syncthreads();
            const int staged_raw = s_nstage;
            SP_MARK(SPS_COLLECT);
            // the list is drained when it is half full (or overflowed, or the walk ends): each drain costs two barriers per
            // 256 survivors plus, now and then, a compaction of the top-k buffer
            if (staged_raw <= SP_STAGE / 2 && blk + 1 < b1) break;
            if (staged_raw == 0) break;
            const int staged = staged_raw < SP_STAGE ? staged_raw : SP_STAGE;
            for (int base = 0; base < staged; base += SP_THREADS) {
                const int i = base + tid;
                const auto tv = tk.view();
                uint64_t h = 0;
                uint32_t l = 0;
                bool have = false;
                if (i < staged) {
                    h = (uint64_t)mono32(stage_sc[i]);
                    l = ~stage_doc[i];
                    have = tk.passes(tv, h, l);
                }
                tk.append(tv, have, h, l);
                tk.settle();
            }
            if (tid == 0) {
                s_nstage = 0;
                if (stats) { s_stat[SPS_STAGED] += staged; s_stat[SPS_RESCANS] += staged_raw > SP_STAGE ? 1 : 0; }
            }
            thr_f = fmaxf(thr_f, tk.threshold_hi32_as_float());
            __syncthreads();
            SP_MARK(SPS_DRAIN);
            if (staged_raw <= SP_STAGE) break;
        }
        if (multi) {                                           // back to group 0 (the prefetched ranges in nb / nl belong to it)
            __syncthreads();
            load_term(0);
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    tk.finalize();
    const int n = tk.count();
    const uint64_t* oh = tk.out_hi();
    const uint32_t* ol = tk.out_lo();
    if (p.n_slices == 1) {
        for (int i = tid; i < p.k; i += SP_THREADS) {
            p.out_scores[(size_t)q * p.k + i] = i < n ? unmono32((uint32_t)oh[i]) : -CUDART_INF_F;
            p.out_ids[(size_t)q * p.k + i] = i < n ? p.id_offset + (int64_t)(~ol[i]) : -1;
        }
        if (tid == 0) p.out_counts[q] = n;
    } else {
        double* ps = p.part_scores + ((size_t)q * p.n_slices + slice) * p.k;
        int64_t* pi = p.part_ids + ((size_t)q * p.n_slices + slice) * p.k;
        for (int i = tid; i < p.k; i += SP_THREADS) {
            ps[i] = i < n ? (double)unmono32((uint32_t)oh[i]) : -CUDART_INF;
            pi[i] = i < n ? p.id_offset + (int64_t)(~ol[i]) : -1;
        }
    }
    SP_MARK(SPS_TOTAL);                                        // init + finalize
    if (stats && tid == 0) {
        const int cta = blockIdx.y * gridDim.x + blockIdx.x;
        if (cta < SP_STAT_CTAS)
            for (int i = 0; i < SP_NSTAT; ++i) p.stats[(size_t)cta * SP_NSTAT + i] = s_stat[i];
    }
#undef SP_MARK
}

static size_t sparse_smem_for(int block_docs, int cap) {
    return (size_t)block_docs * 4 + (size_t)(block_docs / 32) * 4 + (size_t)SP_STAGE * 8 + (size_t)2 * SP_PCAP * 4 +
           (size_t)2 * (SP_PCAP + 32) * 2 + 16 + BlockTopK<SP_THREADS, uint32_t>::smem_bytes(cap) + 64;
}

static size_t sparse_smem(int block_docs, int k, int* cap_out) {
    // streaming top-k buffer: at least k + 512 entries (one compaction per 256 survivors); grown to k + 1024 (one per 768)
    // while two CTAs still fit an SM (each compaction is a multi-pass radix select over the whole buffer)
    int cap = BlockTopK<SP_THREADS, uint32_t>::capacity_for(k, SP_THREADS);
    const int big = k + 1024 > cap ? k + 1024 : cap;
    if (sparse_smem_for(block_docs, big) + 1024 <= 113 * 1024) cap = big;
    if (cap_out) *cap_out = cap;
    return sparse_smem_for(block_docs, cap);
}

static int sparse_slices(int64_t n_blocks, int n_queries) {
    // enough (query, slice) work items to keep two to three CTAs per SM busy and to even out the queries' very different
    // posting counts, but at least four blocks per slice (every slice pays its own top-k warm-up)
    int sm_count = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    int64_t s = (3 * (int64_t)sm_count + n_queries - 1) / n_queries;
    if (s > n_blocks / 4) s = n_blocks / 4;
    const int forced = option(OPT_SPARSE_SLICES, 0);
    if (forced > 0) s = forced;
    if (s > n_blocks) s = n_blocks;
    if (s > SP_MAX_SLICES) s = SP_MAX_SLICES;
    return s < 1 ? 1 : (int)s;
}

}  // namespace b200rag

using namespace b200rag;

extern "C" {

size_t b200rag_sparse_topk_workspace_bytes(int64_t n_docs, int32_t block_docs, int32_t n_queries, int32_t k) {
    if (block_docs <= 0 || n_queries <= 0 || k <= 0) return 0;
    int64_t n_blocks = (n_docs + block_docs - 1) / block_docs;
    if (n_blocks < 1) n_blocks = 1;
    const int64_t s = n_blocks < SP_MAX_SLICES ? n_blocks : SP_MAX_SLICES;
    return align_up((size_t)n_queries * s * k * sizeof(double), 256) + align_up((size_t)n_queries * s * k * sizeof(int64_t), 256) +
           align_up((size_t)n_queries * 4, 256) + 512;
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
    B200_REQUIRE(((uintptr_t)post_doc & 3) == 0 && ((uintptr_t)post_w & 3) == 0, "sparse_topk: postings arrays must be 4-byte aligned");
    B200_REQUIRE(block_docs > 0 && block_docs <= 32768 && block_docs % 32 == 0,
                 "sparse_topk: block_docs must be a multiple of 32 in (0, 32768], got %d", block_docs);
    if (n_queries == 0) return B200RAG_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SparseParams p;
    size_t smem = sparse_smem(block_docs, k, &p.cap);
    int64_t n_blocks = (n_docs + block_docs - 1) / block_docs;
    if (n_blocks < 1) n_blocks = 1;
    B200_REQUIRE(n_queries <= 2147483647 && n_blocks <= 2147483647, "sparse_topk: too many blocks");
    const int ab_flags = option(OPT_SPARSE_FLAGS, 0);
    int n_slices = sparse_slices(n_blocks, n_queries);
    if ((ab_flags & 16) && option(OPT_SPARSE_SLICES, 0) <= 0) {          // round-1 kernel, round-1 slicing
        int sm_count = 148, dev = 0;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
        n_slices = (int)std::max<int64_t>(1, std::min<int64_t>(n_blocks, (2 * (int64_t)sm_count) / n_queries));
    }
    Workspace ws(workspace, workspace_bytes);
    p.part_scores = ws.take<double>((size_t)n_queries * n_slices * k);
    p.part_ids = ws.take<int64_t>((size_t)n_queries * n_slices * k);
    p.gthr = ws.take<unsigned int>((size_t)n_queries);
    if (!ws.ok()) {
        set_error("sparse_topk: workspace too small (%zu < %zu)", workspace_bytes, ws.off);
        return B200RAG_E_WORKSPACE;
    }
    if (n_slices > 1) B200_CUDA_CHECK(cudaMemsetAsync(p.gthr, 0, (size_t)n_queries * 4, st));
    else p.gthr = nullptr;
    p.blk_term_ptr = blk_term_ptr;
    p.post_doc = post_doc;
    p.post_w = post_w;
    p.n_docs = n_docs;
    p.n_terms = n_terms;
    p.block_docs = block_docs;
    p.n_blocks = (int)n_blocks;
    p.n_slices = n_slices;
    p.q_ptr = q_ptr;
    p.q_terms = q_terms;
    p.q_vals = q_vals;
    p.k = k;
    p.id_offset = id_offset;
    p.out_scores = out_scores;
    p.out_ids = out_ids;
    p.out_counts = out_counts;
    p.doc_mask = doc_mask;
    p.flags = ab_flags;
    p.stats = stats_buffer(STATS_SPARSE, (size_t)SP_STAT_CTAS * SP_NSTAT);
    if (p.stats) B200_CUDA_CHECK(cudaMemsetAsync(p.stats, 0, (size_t)SP_STAT_CTAS * SP_NSTAT * 8, st));
    // Queries of up to 15 terms: the mask kernel (sparse_mask.cu), the product path.  Longer queries: the accumulator kernel of
    // this file.  How long the queries are is only known on the device, so both are launched and every CTA of the kernel that
    // does not serve its query leaves at once.
    if (p.flags & 16) {                          // A/B: the round-1 kernel (sparse_r1.cu) for every query
        int rc = launch_sparse_r1(p, n_queries, st);
        if (rc) return rc;
        return launch_merge(p.part_scores, p.part_ids, n_queries, nullptr, n_slices * k, k, nullptr, out_scores, out_ids, out_counts, st);
    }
    if (smem > 225 * 1024) {
        set_error("sparse_topk: block_docs=%d with k=%d needs %zu bytes of shared memory (max 230400); use a smaller block",
                  block_docs, k, smem);
        return B200RAG_E_UNSUPPORTED;
    }
    const size_t mask_smem = sparse_mask_smem(block_docs, k, nullptr);
    const bool use_mask = !(p.flags & 8) && mask_smem <= 225 * 1024 && ((uintptr_t)post_doc & 15) == 0 && ((uintptr_t)post_w & 15) == 0 &&
                          block_docs % 8 == 0;
    p.mask_max_terms = use_mask ? SPM_MAX_TERMS : -1;
    if (use_mask) {
        int rc = launch_sparse_mask(p, n_queries, st);
        if (rc) return rc;
    }
    B200_CUDA_CHECK(cudaFuncSetAttribute(sparse_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)n_queries, (unsigned)n_slices);
    sparse_query_kernel<<<grid, SP_THREADS, smem, st>>>(p); count_launch();
    B200_CUDA_CHECK(cudaGetLastError());
    if (n_slices == 1) return B200RAG_OK;
    return launch_merge(p.part_scores, p.part_ids, n_queries, nullptr, n_slices * k, k, nullptr, out_scores, out_ids, out_counts, st);
}

}  // extern "C"
