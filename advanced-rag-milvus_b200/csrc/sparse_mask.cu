// sparse_mask.cu -- sparse inner-product (BM25-weighted) top-k over doc-range-blocked postings, queries of up to 15 terms.
// EXPERIMENTAL (round 2): selectable with option "sparse_flags" bit 3 (8) for A/B runs; NOT the product path -- on the same box it
// measured 1.2 - 1.4 ms at config 4 against 0.65 ms for sparse_bm25.cu (321 M warp instructions against 273 M: the per-block
// bookkeeping of 8192-document blocks outweighs what the missing accumulator traffic saves).  Results are bit-identical.
//
// No per-document accumulators.  A document's score is a short chain  s = fmaf(qv_t, w_t, s)  over the query terms it holds, in
// ascending term id (the canonical order, bit-identical to oracle/exact_scan.c:orc_sparse_topk); most documents a query touches
// hold exactly ONE of its terms.  So instead of read-modify-writing accumulators list after list (round 1: a block barrier
// between terms, 64 KB of accumulators scanned per block, 7.9 warp instructions per posting), a block is processed as two flat,
// order-free passes over its postings, every thread busy in every iteration:
//
//   mark    for posting (term j, row d):  atomicOr(mask[d], 1 << j)  -- a 16-bit term mask per document of the block;
//   score   for posting (j, d): the posting whose term is the LOWEST bit of mask[d] owns the document.  The owner starts the
//           chain with its own weight and, for every further bit of the mask (rare), finds the document in that term's
//           doc-sorted list by binary search and continues the chain -- all in registers.  A score that reaches the running
//           k-th best `thr` goes to the survivor list; bit 15 of the mask records "scored" so that a re-run of the pass (the list
//           overflowed, was drained, thr went up) skips finished documents.  A document whose score is exactly 0 is still a hit
//           while thr <= 0: ownership comes from the mask, not from a non-zero accumulator, so there is one collect path.
//   the masks are wiped with 16-byte stores afterwards (32 KB per 16384-document block).
//
// The postings of the NEXT block are staged in shared memory with 16-byte cp.async copies while the current one is processed
// (ranges are fetched two blocks ahead; each warp copies two of the up-to-15 lists), so neither pass waits for global memory.
// Blocks whose lists do not fit the staging buffer read their postings in place -- same code, same results.  Survivors drain into
// the CTA's streaming top-k (select.cuh) when their list is half full.  One 256-thread CTA per (query, slice of blocks), two CTAs
// per SM; slices of one query share thresholds through a global atomicMax; merge_topk_kernel reduces the slices.
// Algorithmic HBM traffic = 6 bytes per posting of the query's terms.
#include "sparse.cuh"

namespace b200rag {

constexpr int SPM_THREADS = 256;
constexpr int SPM_WARPS = SPM_THREADS / 32;
constexpr int SPM_TG = 16;           // list slots (15 usable terms: bit 15 of the mask is the "scored" flag)
constexpr int SPM_STAGE = 1024;      // survivor list entries
constexpr int SPM_PCAP = 3072;       // postings of one block staged per buffer
constexpr int SPM_WPAD = 6 * SPM_TG, SPM_DPAD = 14 * SPM_TG;      // alignment slack of the 16-byte copies, per buffer
constexpr int SPM_RING = 3;
constexpr int SPM_NSTAT = 12;
constexpr int SPM_STAT_CTAS = 1024;
constexpr uint32_t SPM_DONE = 0x8000u;
constexpr int SPM_MULTI = 1024;      // queued multi-term documents per score pass (more are chained inline)

enum { SMS_TOTAL = 0, SMS_MARK, SMS_SCORE, SMS_DRAIN, SMS_BLOCKS, SMS_STAGED, SMS_RESCANS, SMS_MULTI, SMS_POSTINGS, SMS_WAIT, SMS_UNSTAGED };

__device__ __forceinline__ void spm_cp_async16(void* smem_dst, const void* gsrc, unsigned src_bytes) {
    // 16-byte asynchronous global -> shared copy; src_bytes < 16 zero-fills the rest (nothing beyond src_bytes is read)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc),
                 "r"(src_bytes)
                 : "memory");
}

struct SpmSlot {                     // one block's lists: ranges in the index and where their staged copies live
    long long beg[SPM_TG];
    int len[SPM_TG];
    int off[SPM_TG + 1];             // exclusive prefix of len: the flat index space of the two passes
    int wbase[SPM_TG];               // staged weight of posting i of list j: sw[wbase[j] + i]
    int dbase[SPM_TG];               // staged row    of posting i of list j: sd[dbase[j] + i]
    int staged;                      // every list of the block is staged (all or nothing)
};

__global__ void __launch_bounds__(SPM_THREADS, 2) sparse_mask_kernel(const SparseParams p) {
    extern __shared__ __align__(16) char smem[];
    __shared__ SpmSlot s_slot[SPM_RING];
    __shared__ float s_qv[SPM_TG];
    __shared__ int s_nstage, s_nmulti;
    __shared__ uint32_t multi[SPM_MULTI];            // owners of multi-term documents of the current block: (list << 16) | posting
    __shared__ unsigned int s_gthr;
    __shared__ unsigned long long s_stat[SPM_NSTAT];
    __shared__ long long s_last;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q = blockIdx.x, slice = blockIdx.y;
    const int64_t qs = p.q_ptr[q];
    const int nq = (int)(p.q_ptr[q + 1] - qs);
    if (nq > p.mask_max_terms) return;                          // served by sparse_query_kernel

    const int block_docs = p.block_docs;
    uint16_t* mask16 = reinterpret_cast<uint16_t*>(smem);                               // [block_docs] term masks
    uint32_t* mask32 = reinterpret_cast<uint32_t*>(smem);
    uint32_t* stage_doc = mask32 + block_docs / 2;                                      // [SPM_STAGE] row inside this shard
    float* stage_sc = reinterpret_cast<float*>(stage_doc + SPM_STAGE);                  // [SPM_STAGE]
    float* sw = stage_sc + SPM_STAGE;                                                   // [2][SPM_PCAP + SPM_WPAD] staged weights
    constexpr int SW_STRIDE = SPM_PCAP + SPM_WPAD, SD_STRIDE = SPM_PCAP + SPM_DPAD;
    uint16_t* sd = reinterpret_cast<uint16_t*>(sw + 2 * SW_STRIDE);                     // [2][SPM_PCAP + SPM_DPAD] staged rows
    char* tkmem = reinterpret_cast<char*>(sd + 2 * SD_STRIDE);
    tkmem = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(tkmem) + 15) & ~uintptr_t(15));
    using TopK = BlockTopK<SPM_THREADS, uint32_t>;
    TopK tk;
    tk.attach(tkmem, p.cap, p.k, SPM_THREADS, /*start_digit=*/TopK::NLO + 3);
    tk.init();
    {
        uint4* m4 = reinterpret_cast<uint4*>(smem);
        for (int i = tid; i < block_docs / 8; i += SPM_THREADS) m4[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    const bool stats = p.stats != nullptr;
    if (tid == 0) {
        s_nstage = 0;
        s_gthr = 0u;
        if (stats) {
            for (int i = 0; i < SPM_NSTAT; ++i) s_stat[i] = 0;
            s_last = clock64();
        }
    }
#define SPM_MARK(i)                                                      \
    do {                                                                 \
        if (stats && tid == 0) {                                         \
            const long long now_ = clock64();                            \
            s_stat[i] += (unsigned long long)(now_ - s_last);            \
            s_last = now_;                                               \
        }                                                                \
    } while (0)

    const int b0 = (int)((int64_t)slice * p.n_blocks / p.n_slices), b1 = (int)((int64_t)(slice + 1) * p.n_blocks / p.n_slices);
    const size_t row_stride = (size_t)p.n_terms + 1;
    // postings in the whole index = the end pointer of the last block (bounds the 16-byte copies at the very end)
    const long long nnz = p.blk_term_ptr[(size_t)(p.n_blocks - 1) * row_stride + p.n_terms];
    const bool staging = !(p.flags & 2);

    // lane j < SPM_TG of warp 0 owns term j of the query
    int my_t = -1;
    if (tid < SPM_TG) {
        float qv = 0.f;
        if (tid < nq) {
            const int t = p.q_terms[qs + tid];
            if (t >= 0 && t < p.n_terms) { my_t = t; qv = p.q_vals[qs + tid]; }
        }
        s_qv[tid] = qv;
    }
    auto fetch_range = [&](int blk, long long& rb, int& rl) {
        rb = 0;
        rl = 0;
        if (my_t >= 0) {
            const int64_t* src = p.blk_term_ptr + (size_t)blk * row_stride + my_t;
            rb = src[0];
            rl = (int)(src[1] - rb);
        }
    };
    // Warp 0 publishes the ranges its lanes hold into a ring slot together with the staging layout; followed by a barrier.
    // 16-byte copies start at 16-byte boundaries of the SOURCE, so a list keeps its offset inside its first chunk.
    auto publish = [&](int slot, long long rb, int rl) {
        if (warp != 0) return;
        const int len = tid < SPM_TG ? rl : 0;
        int incl = len;
#pragma unroll
        for (int o = 1; o < SPM_TG; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        const int total = __shfl_sync(0xffffffffu, incl, SPM_TG - 1);
        const bool st = staging && total <= SPM_PCAP;
        const int wsh = (int)(rb & 3), dsh = (int)(rb & 7);
        const int wch = (st && len) ? ((wsh + len + 3) >> 2) : 0, dch = (st && len) ? ((dsh + len + 7) >> 3) : 0;
        int wincl = wch, dincl = dch;
#pragma unroll
        for (int o = 1; o < SPM_TG; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, wincl, o), u = __shfl_up_sync(0xffffffffu, dincl, o);
            if (lane >= o) { wincl += v; dincl += u; }
        }
        if (tid < SPM_TG) {
            SpmSlot& s = s_slot[slot];
            s.beg[tid] = rb;
            s.len[tid] = len;
            s.off[tid] = incl - len;
            s.wbase[tid] = 4 * (wincl - wch) + wsh;
            s.dbase[tid] = 8 * (dincl - dch) + dsh;
            if (tid == SPM_TG - 1) { s.off[SPM_TG] = incl; s.staged = st ? 1 : 0; }
        }
    };
    // All threads: start the asynchronous copies of a published block into staging buffer `buf` (warp w: lists w, w + 8).
    auto issue_stage = [&](int slot, int buf) {
        const SpmSlot& s = s_slot[slot];
        if (!s.staged) return;
        float* wdst = sw + buf * SW_STRIDE;
        uint16_t* ddst = sd + buf * SD_STRIDE;
        for (int j = warp; j < SPM_TG; j += SPM_WARPS) {
            const int len = s.len[j];
            if (len == 0) continue;
            const long long beg = s.beg[j];
            const int wsh = (int)(beg & 3), dsh = (int)(beg & 7);
            const long long wg0 = beg - wsh, dg0 = beg - dsh;                  // 16-byte aligned posting indices
            const int wch = (wsh + len + 3) >> 2, dch = (dsh + len + 7) >> 3;
            float* wd = wdst + (s.wbase[j] - wsh);
            uint16_t* dd = ddst + (s.dbase[j] - dsh);
            for (int c = lane; c < wch; c += 32) {
                const long long g = wg0 + 4 * (long long)c;
                const long long left = nnz - g;
                spm_cp_async16(wd + 4 * c, p.post_w + g, left >= 4 ? 16u : (unsigned)(left * 4));
            }
            for (int c = lane; c < dch; c += 32) {
                const long long g = dg0 + 8 * (long long)c;
                const long long left = nnz - g;
                spm_cp_async16(dd + 8 * c, p.post_doc + g, left >= 8 ? 16u : (unsigned)(left * 2));
            }
        }
    };

    // a survivor goes to the list; false = the list is full (the document stays unscored for the re-run of the pass)
    auto stage = [&](uint32_t doc, float sc) -> bool {
        const int slot = atomicAdd(&s_nstage, 1);
        if (slot >= SPM_STAGE) return false;
        stage_doc[slot] = doc;
        stage_sc[slot] = sc;
        return true;
    };

    // ---- prologue: ranges of the first block published and staged, ranges of the second one in flight
    long long nb = 0;
    int nl = 0;
    const bool walk = nq > 0 && b0 < b1;
    if (walk) fetch_range(b0, nb, nl);
    __syncthreads();                                           // (masks wiped, top-k initialised, s_qv written)
    if (walk) {
        publish(b0 % SPM_RING, nb, nl);
        __syncthreads();
        issue_stage(b0 % SPM_RING, b0 & 1);
        if (b0 + 1 < b1) fetch_range(b0 + 1, nb, nl);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    SPM_MARK(SMS_TOTAL);
    float thr_f = -CUDART_INF_F;
    for (int blk = b0; blk < b1 && nq > 0; ++blk) {
        const int64_t doc0 = (int64_t)blk * block_docs;
        const int buf = blk & 1;
        const SpmSlot& s = s_slot[blk % SPM_RING];
        // ---- block top: publish + stage the NEXT block, exchange thresholds with the other slices, wait for THIS block's copies
        if (blk + 1 < b1) publish((blk + 1) % SPM_RING, nb, nl);
        if (tid == 0 && p.gthr) {
            if (tk.st->has_thr) atomicMax(p.gthr + q, (unsigned int)tk.st->thr_hi);
            unsigned int g;
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(g) : "l"(p.gthr + q) : "memory");
            s_gthr = g;
        }
        __syncthreads();
        if (blk + 1 < b1) issue_stage((blk + 1) % SPM_RING, buf ^ 1);
        asm volatile("cp.async.commit_group;" ::: "memory");   // (one group per block, possibly empty)
        if (blk + 2 < b1) fetch_range(blk + 2, nb, nl);        // (in flight during this whole block)
        asm volatile("cp.async.wait_group 1;" ::: "memory");   // everything but the newest group has landed: this block's
        __syncthreads();
        SPM_MARK(SMS_WAIT);
        if (s_gthr) thr_f = fmaxf(thr_f, unmono32(s_gthr));
        const int total = s.off[SPM_TG];
        const bool staged = s.staged != 0;                      // CTA-uniform
        const float* swb = sw + buf * SW_STRIDE;
        const uint16_t* sdb = sd + buf * SD_STRIDE;
        auto row_at = [&](int j, int i) -> int {
            return staged ? (int)sdb[s.dbase[j] + i] : (int)__ldg(p.post_doc + s.beg[j] + i);
        };
        auto w_at = [&](int j, int i) -> float {
            return staged ? swb[s.wbase[j] + i] : __ldg(p.post_w + s.beg[j] + i);
        };
        // the canonical chain of a document whose lowest term is list j (posting i): its own weight, then every further term
        // of its mask in ascending order, each found by binary search in that term's row-sorted list
        auto chain_score = [&](int j, int i, int d, uint32_t m) -> float {
            float sc = fmaf(s_qv[j], w_at(j, i), 0.0f);
            uint32_t rest = (m & ~SPM_DONE) >> (j + 1);
            int j2 = j + 1;
            while (rest) {
                const int sk = __ffs((int)rest) - 1;
                j2 += sk;
                rest >>= sk + 1;
                int lo = 0, hi = s.len[j2];
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (row_at(j2, mid) < d) lo = mid + 1;
                    else hi = mid;
                }
                sc = fmaf(s_qv[j2], w_at(j2, lo), sc);
                ++j2;
            }
            return sc;
        };
        // ---- mark: which of the query's terms does every document of the block hold?
        {
            int j = 0;
            for (int pos = tid; pos < total; pos += SPM_THREADS) {
                while (pos >= s.off[j + 1]) ++j;
                const int d = row_at(j, pos - s.off[j]);
                atomicOr(&mask32[d >> 1], (1u << j) << ((d & 1) * 16));
            }
        }
        __syncthreads();
        SPM_MARK(SMS_MARK);
        if (stats && tid == 0) {
            s_stat[SMS_BLOCKS] += 1;
            s_stat[SMS_POSTINGS] += (unsigned long long)total;
            s_stat[SMS_UNSTAGED] += staged ? 0 : 1;
        }
        // ---- score (repeated while the survivor list overflows).  Documents with ONE query term -- the great majority -- are
        //      scored on the spot; owners of documents with several terms are queued and handled afterwards with every lane of
        //      a warp busy on a queued document (a binary search per further term: done inline it would stall the 31 other lanes
        //      of nearly every warp iteration).
        for (;;) {
            if (tid == 0) s_nmulti = 0;
            __syncthreads();
            auto finish_doc = [&](int d, uint32_t m, float sc) {
                bool keep = !(sc < thr_f);
                if (keep && p.doc_mask) {
                    const int64_t g = doc0 + d;
                    keep = (__ldg(p.doc_mask + (g >> 5)) >> (g & 31)) & 1u;
                }
                if (!keep || stage((uint32_t)(doc0 + d), sc)) mask16[d] = (uint16_t)(m | SPM_DONE);   // only the owner writes
            };
            {
                int j = 0;
                for (int pos = tid; pos < total; pos += SPM_THREADS) {
                    while (pos >= s.off[j + 1]) ++j;
                    const int i = pos - s.off[j];
                    const int d = row_at(j, i);
                    const uint32_t m = mask16[d];
                    if (m & (SPM_DONE | ((1u << j) - 1u))) continue;     // a lower term owns the document, or it is finished
                    if (m >> (j + 1)) {                                  // further terms: queue (j, i) for the second phase
                        const int at = atomicAdd(&s_nmulti, 1);
                        if (at < SPM_MULTI) multi[at] = ((uint32_t)j << 16) | (uint32_t)i;
                        else finish_doc(d, m, chain_score(j, i, d, m));  // (queue full: inline after all)
                    } else {
                        finish_doc(d, m, fmaf(s_qv[j], w_at(j, i), 0.0f));
                    }
                }
            }
            __syncthreads();
            {
                const int n_multi = min(s_nmulti, SPM_MULTI);
                for (int idx = tid; idx < n_multi; idx += SPM_THREADS) {
                    const uint32_t e = multi[idx];
                    const int j = (int)(e >> 16), i = (int)(e & 0xffffu);
                    const int d = row_at(j, i);
                    const uint32_t m = mask16[d];
                    finish_doc(d, m, chain_score(j, i, d, m));
                }
                if (stats && tid == 0) s_stat[SMS_MULTI] += (unsigned long long)s_nmulti;
            }
            __syncthreads();
            const int staged_raw = s_nstage;
            SPM_MARK(SMS_SCORE);
            // the list is drained when it is half full (or overflowed, or the walk ends)
            if (staged_raw <= SPM_STAGE / 2 && blk + 1 < b1) break;
            if (staged_raw == 0) break;
            const int n_st = staged_raw < SPM_STAGE ? staged_raw : SPM_STAGE;
            for (int base = 0; base < n_st; base += SPM_THREADS) {
                const int i = base + tid;
                const auto tv = tk.view();
                uint64_t h = 0;
                uint32_t l = 0;
                bool have = false;
                if (i < n_st) {
                    h = (uint64_t)mono32(stage_sc[i]);
                    l = ~stage_doc[i];
                    have = tk.passes(tv, h, l);
                }
                tk.append(tv, have, h, l);
                tk.settle();
            }
            if (tid == 0) {
                s_nstage = 0;
                if (stats) { s_stat[SMS_STAGED] += n_st; s_stat[SMS_RESCANS] += staged_raw > SPM_STAGE ? 1 : 0; }
            }
            thr_f = fmaxf(thr_f, tk.threshold_hi32_as_float());
            __syncthreads();
            SPM_MARK(SMS_DRAIN);
            if (staged_raw <= SPM_STAGE) break;
        }
        // ---- wipe the masks (the next block's mark pass starts behind the barrier at its top)
        {
            uint4* m4 = reinterpret_cast<uint4*>(smem);
            for (int i = tid; i < block_docs / 8; i += SPM_THREADS) m4[i] = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    tk.finalize();
    const int n = tk.count();
    const uint64_t* oh = tk.out_hi();
    const uint32_t* ol = tk.out_lo();
    {   // per-slice result; merge_topk_kernel reduces the slices (and converts back to fp32)
        double* ps = p.part_scores + ((size_t)q * p.n_slices + slice) * p.k;
        int64_t* pi = p.part_ids + ((size_t)q * p.n_slices + slice) * p.k;
        for (int i = tid; i < p.k; i += SPM_THREADS) {
            ps[i] = i < n ? (double)unmono32((uint32_t)oh[i]) : -CUDART_INF;
            pi[i] = i < n ? p.id_offset + (int64_t)(~ol[i]) : -1;
        }
    }
    SPM_MARK(SMS_TOTAL);                                       // init + finalize
    if (stats && tid == 0) {
        const int cta = blockIdx.y * gridDim.x + blockIdx.x;
        if (cta < SPM_STAT_CTAS)
            for (int i = 0; i < SPM_NSTAT; ++i) p.stats[(size_t)cta * SPM_NSTAT + i] = s_stat[i];
    }
#undef SPM_MARK
}

static size_t sparse_mask_smem_for(int block_docs, int cap) {
    return (size_t)block_docs * 2 + (size_t)SPM_STAGE * 8 + (size_t)2 * (SPM_PCAP + SPM_WPAD) * 4 +
           (size_t)2 * (SPM_PCAP + SPM_DPAD) * 2 + 16 + BlockTopK<SPM_THREADS, uint32_t>::smem_bytes(cap) + 64;
}

size_t sparse_mask_smem(int block_docs, int k, int* cap_out) {
    // streaming top-k buffer: at least k + 512 entries (one compaction per 256 survivors); grown to k + 1024 (one per 768)
    // while two CTAs still fit an SM (each compaction is a multi-pass radix select over the whole buffer)
    int cap = BlockTopK<SPM_THREADS, uint32_t>::capacity_for(k, SPM_THREADS);
    const int big = k + 1024 > cap ? k + 1024 : cap;
    if (sparse_mask_smem_for(block_docs, big) + 2048 <= 113 * 1024) cap = big;
    if (cap_out) *cap_out = cap;
    return sparse_mask_smem_for(block_docs, cap);
}

int launch_sparse_mask(const SparseParams& p, int n_queries, cudaStream_t st) {
    int cap = 0;
    const size_t smem = sparse_mask_smem(p.block_docs, p.k, &cap);
    SparseParams pm = p;
    pm.cap = cap;
    B200_CUDA_CHECK(cudaFuncSetAttribute(sparse_mask_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)n_queries, (unsigned)p.n_slices);
    sparse_mask_kernel<<<grid, SPM_THREADS, smem, st>>>(pm); count_launch();
    B200_CUDA_CHECK(cudaGetLastError());
    return B200RAG_OK;
}

}  // namespace b200rag
