// select.cuh -- block-level streaming top-k over 96/128-bit keys held in shared memory.
//
// Key = (hi: u64, lo: u32 or u64), ordered lexicographically; "greater is better".  Callers encode
//   hi = order-preserving image of the score (mono32 / mono64), lo = ~local_id
// so that "greater" == (score descending, id ascending): the ranking rule of b200rag.h.
//
// Usage (all threads of the block, in lock step):
//   tk.attach(...); tk.init(); __syncthreads();
//   loop { tk.offer(valid, hi, lo); [more offers]; tk.settle(); }   // settle() = barrier + compaction when nearly full;
//                                                                   // offer() is called by whole warps (valid = false to skip)
//   __syncthreads(); tk.finalize();   // <= k entries, sorted best first, in tk.out_hi()/out_lo(), count in tk.count()
//
// Capacity contract: cap >= k + reserve and cap >= next_pow2(k), where `reserve` bounds the number of offers the
// whole block makes between two settle() calls.
#pragma once
#include "common.cuh"

namespace b200rag {

struct TopKShared {
    int count;      // entries in the active buffer
    int active;     // which ping-pong buffer is live
    int has_thr;    // thr_* valid: entries <= thr were already discarded and may be refused
    int sel_digit;
    int sel_want;
    int sel_exact;  // the selected bucket holds exactly sel_want entries: the remaining digits cannot matter
    int out_count;
    int eq_taken;
    int pad;
    unsigned long long thr_hi;
    unsigned long long thr_lo;
    int hist[256];
};

template <typename LoT>
__device__ __forceinline__ bool key_gt(uint64_t ah, LoT al, uint64_t bh, LoT bl) {
    return ah > bh || (ah == bh && al > bl);
}

__host__ __device__ inline int next_pow2_int(int x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

template <int THREADS, typename LoT = uint32_t>
struct BlockTopK {
    static constexpr int NLO = (int)sizeof(LoT);        // key bytes contributed by lo
    static constexpr int TOP_DIGIT = 8 + NLO - 1;
    TopKShared* st;
    uint64_t* hi[2];
    LoT* lo[2];
    int cap;
    int k;
    int reserve;      // max offers (whole block) between two settle() calls
    int start_digit;  // highest key byte that can differ (TOP_DIGIT = whole key; NLO+3 if hi only uses its low 32 bits)

    __host__ __device__ static int capacity_for(int k, int reserve) {
        int c = k + 2 * reserve;
        int p = next_pow2_int(k);
        return c > p ? c : p;
    }
    __host__ __device__ static size_t smem_bytes(int cap) {
        return sizeof(TopKShared) + (size_t)cap * (8 + sizeof(LoT)) * 2 + 32;
    }
    // Carve from a 16-byte aligned smem pointer; returns the pointer past the carved region.
    __device__ char* attach(char* p, int cap_, int k_, int reserve_, int start_digit_ = TOP_DIGIT) {
        cap = cap_;
        k = k_;
        reserve = reserve_;
        start_digit = start_digit_;
        st = reinterpret_cast<TopKShared*>(p);
        p += sizeof(TopKShared);
        hi[0] = reinterpret_cast<uint64_t*>(p); p += (size_t)cap * 8;
        hi[1] = reinterpret_cast<uint64_t*>(p); p += (size_t)cap * 8;
        lo[0] = reinterpret_cast<LoT*>(p); p += (size_t)cap * sizeof(LoT);
        lo[1] = reinterpret_cast<LoT*>(p); p += (size_t)cap * sizeof(LoT);
        return reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(p) + 15) & ~uintptr_t(15));
    }
    __device__ void init() {
        if (threadIdx.x == 0) {
            st->count = 0;
            st->active = 0;
            st->has_thr = 0;
            st->thr_hi = 0;
            st->thr_lo = 0;
        }
    }
    // Whole warps call this together (converged; `valid` false for lanes with nothing to offer): the accepted lanes of a
    // warp take their slots with one atomic.
    __device__ __forceinline__ void offer(bool valid, uint64_t h, LoT l) {
        const bool pass = valid && (!st->has_thr || key_gt<LoT>(h, l, st->thr_hi, (LoT)st->thr_lo));
        const unsigned bal = __ballot_sync(0xffffffffu, pass);
        if (bal) {
            const int lane = threadIdx.x & 31, leader = __ffs((int)bal) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(&st->count, __popc(bal));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (pass) {
                const int i = base + __popc(bal & ((1u << lane) - 1u));
                const int a = st->active;
                hi[a][i] = h;
                lo[a][i] = l;
            }
        }
    }
    // Register copy of what offer() reads, for loops that make many offers between two settles (nothing changes meanwhile).
    struct View {
        int has_thr, active;
        uint64_t thr_hi;
        LoT thr_lo;
    };
    __device__ __forceinline__ View view() const {
        View v;
        v.has_thr = st->has_thr;
        v.active = st->active;
        v.thr_hi = st->thr_hi;
        v.thr_lo = (LoT)st->thr_lo;
        return v;
    }
    __device__ __forceinline__ bool passes(const View& v, uint64_t h, LoT l) const {
        return !v.has_thr || key_gt<LoT>(h, l, v.thr_hi, v.thr_lo);
    }
    // offer() against a view: whole warps call it together, `have` = this lane holds a key that passes(v, ...).
    // Returns false when no lane of the warp had one (the warp-uniform exit of an append loop).
    __device__ __forceinline__ bool append(const View& v, bool have, uint64_t h, LoT l) {
        const unsigned bal = __ballot_sync(0xffffffffu, have);
        if (!bal) return false;
        const int lane = threadIdx.x & 31, leader = __ffs((int)bal) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(&st->count, __popc(bal));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (have) {
            const int i = base + __popc(bal & ((1u << lane) - 1u));
            hi[v.active][i] = h;
            lo[v.active][i] = l;
        }
        return true;
    }
    // Bulk form for a warp that knows how many keys it is going to add: lane 0 takes `n_warp` consecutive slots (warp-uniform
    // argument, whole warp calls), every lane then put()s its keys at first + (its offset inside the warp's range).
    // Keys at or below the threshold may be put as well: the next compaction drops them.
    __device__ __forceinline__ int reserve_warp(int n_warp) {
        int base = 0;
        if ((threadIdx.x & 31) == 0) base = atomicAdd(&st->count, n_warp);
        return __shfl_sync(0xffffffffu, base, 0);
    }
    __device__ __forceinline__ void put(const View& v, int i, uint64_t h, LoT l) {
        hi[v.active][i] = h;
        lo[v.active][i] = l;
    }
    // Would offer() accept this key?  (Lets callers skip losers before spending one of their `reserve` offers.)
    __device__ __forceinline__ bool passes(uint64_t h, LoT l) const {
        return !st->has_thr || key_gt<LoT>(h, l, st->thr_hi, (LoT)st->thr_lo);
    }
    // For keys whose hi word is mono32(score): the score of the current k-th best, or -inf while fewer than k were seen.
    // Anything strictly below it can be dropped without building a key.
    __device__ __forceinline__ float threshold_hi32_as_float() const {
        return st->has_thr ? unmono32((uint32_t)st->thr_hi) : -CUDART_INF_F;
    }
    // Barrier + (when the next round could overflow) compaction.  All threads must call.  Ends synchronised.
    __device__ __forceinline__ void settle() {
        __syncthreads();
        const int c = st->count;
        __syncthreads();          // nobody may append again before everyone has read the count
        if (c > cap - reserve) compact();
    }
    __device__ int count() const { return st->count; }
    __device__ const uint64_t* out_hi() const { return hi[st->active]; }
    __device__ const LoT* out_lo() const { return lo[st->active]; }

    __device__ static __forceinline__ int key_digit(uint64_t h, LoT l, int d) {
        return d >= NLO ? (int)((h >> ((d - NLO) * 8)) & 255) : (int)((l >> (d * 8)) & 255);
    }

    // Keep the k greatest entries (precondition: all threads call; entry/exit synchronised).
    // exact = false (streaming use): up to reserve/2 further entries may stay -- the radix walk stops at the first digit
    // whose boundary bucket overshoots k by no more than that.  Result lists are full of tied scores (equal BM25 weights,
    // duplicated rows), and splitting a tie group exactly costs a walk over all the id digits as well; the final
    // compaction (finalize) does that once.
    __device__ void compact(bool exact = false) {
        const int tid = threadIdx.x;
        const int n = st->count;
        const int a = st->active;
        __syncthreads();   // everyone has read count/active before anyone mutates
        if (n <= k) return;
        const uint64_t* bh = hi[a];
        const LoT* bl = lo[a];
        uint64_t sel_hi = 0, mask_hi = 0;
        LoT sel_lo = 0, mask_lo = 0;
        int want = k;
        const int slack = exact ? 0 : reserve / 2;
        if (start_digit < TOP_DIGIT) {   // bytes above start_digit are identical for all entries: part of the prefix from the start
            mask_hi = ~uint64_t(0) << ((start_digit - NLO + 1) * 8);
            sel_hi = bh[0] & mask_hi;
        }
        bool early = false;              // stopped before the last digit: every entry with the selected prefix is kept
        for (int d = start_digit; d >= 0; --d) {
            for (int i = tid; i < 256; i += THREADS) st->hist[i] = 0;
            __syncthreads();
            // Scores of one result list share their top bytes, so the first passes put (nearly) every entry into one bin:
            // lanes with equal bins are counted with a single atomic.
            for (int i0 = 0; i0 < n; i0 += THREADS) {
                const int i = i0 + tid;
                bool ok = false;
                int bin = 0;
                if (i < n) {
                    const uint64_t h = bh[i];
                    const LoT l = bl[i];
                    ok = (h & mask_hi) == sel_hi && (l & mask_lo) == sel_lo;
                    bin = key_digit(h, l, d);
                }
                const unsigned act = __ballot_sync(0xffffffffu, ok);
                if (ok) {
                    const unsigned peers = __match_any_sync(act, bin);
                    if ((int)(tid & 31) == __ffs((int)peers) - 1) atomicAdd(&st->hist[bin], __popc(peers));
                }
            }
            __syncthreads();
            if (tid < 32) {
                // lane j owns bins [8j, 8j+8); find the highest bin b with  #(bins > b) < want <= #(bins >= b)
                int c[8];
                int s = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) { c[j] = st->hist[tid * 8 + j]; s += c[j]; }
                int incl = s;   // inclusive suffix sum over lanes >= tid
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    int v = __shfl_down_sync(0xffffffffu, incl, off);
                    if (tid + off < 32) incl += v;
                }
                int above = incl - s;   // entries in lanes > tid
                if (above < want && want <= incl) {
                    int cum = above;
                    int b = 7;
                    for (; b > 0; --b) {
                        if (cum + c[b] >= want) break;
                        cum += c[b];
                    }
                    st->sel_digit = tid * 8 + b;
                    st->sel_want = want - cum;
                    st->sel_exact = c[b] - (want - cum) <= slack;
                }
            }
            __syncthreads();
            const int b = st->sel_digit;
            want = st->sel_want;
            if (d >= NLO) {
                sel_hi |= (uint64_t)b << ((d - NLO) * 8);
                mask_hi |= (uint64_t)255 << ((d - NLO) * 8);
            } else {
                sel_lo |= (LoT)b << (d * 8);
                mask_lo |= (LoT)255 << (d * 8);
            }
            if (st->sel_exact && d > 0) {   // (no ties at the boundary is the common case: half of the passes or more are skipped)
                early = true;
                break;
            }
        }
        // complete walk: (sel_hi, sel_lo) is the k-th greatest key and `want` of the entries equal to it are kept.
        // early stop:    entries whose masked key is >= the selected prefix are kept (k of them, plus at most `slack`).
        if (tid == 0) { st->out_count = 0; st->eq_taken = 0; }
        __syncthreads();
        uint64_t* oh = hi[a ^ 1];
        LoT* ol = lo[a ^ 1];
        for (int i0 = 0; i0 < n; i0 += THREADS) {
            const int i = i0 + tid;
            uint64_t h = 0;
            LoT l = 0;
            bool keep = false;
            if (i < n) {
                h = bh[i];
                l = bl[i];
                if (early) {
                    const uint64_t mh = h & mask_hi;
                    const LoT ml = l & mask_lo;
                    keep = mh > sel_hi || (mh == sel_hi && ml >= sel_lo);
                } else {
                    keep = key_gt<LoT>(h, l, sel_hi, sel_lo);
                    if (!keep && h == sel_hi && l == sel_lo) keep = atomicAdd(&st->eq_taken, 1) < want;
                }
            }
            // one slot reservation per warp
            const unsigned kept = __ballot_sync(0xffffffffu, keep);
            if (kept) {
                const int lane = tid & 31, leader = __ffs((int)kept) - 1;
                int base = 0;
                if (lane == leader) base = atomicAdd(&st->out_count, __popc(kept));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (keep) {
                    const int j = base + __popc(kept & ((1u << lane) - 1u));
                    oh[j] = h;
                    ol[j] = l;
                }
            }
        }
        if (early) {
            // threshold for later offers: the greatest key below every kept one = (prefix, low bits 0) - 1
            if (sel_lo != 0) { sel_lo = (LoT)(sel_lo - 1); }
            else { sel_hi -= 1; sel_lo = (LoT)~(LoT)0; }     // sel_hi > 0: a zero prefix would have kept all n > k entries
        }
        __syncthreads();
        if (tid == 0) {
            st->count = st->out_count;
            st->active = a ^ 1;
            st->has_thr = 1;
            st->thr_hi = sel_hi;
            st->thr_lo = sel_lo;
        }
        __syncthreads();
    }

    // Reduce to <= k entries and sort them best first.  Entry: synchronised.  Exit: synchronised.
    __device__ void finalize() {
        compact(/*exact=*/true);
        const int tid = threadIdx.x;
        const int a = st->active;
        const int n = st->count;
        uint64_t* bh = hi[a];
        LoT* bl = lo[a];
        const int n2 = next_pow2_int(n > 1 ? n : 1);
        for (int i = n + tid; i < n2; i += THREADS) { bh[i] = 0; bl[i] = 0; }
        __syncthreads();
        for (int size = 2; size <= n2; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                for (int t = tid; t < n2 / 2; t += THREADS) {
                    int i = 2 * t - (t & (stride - 1));
                    int j = i + stride;
                    bool desc = (i & size) == 0;
                    uint64_t hi_i = bh[i], hi_j = bh[j];
                    LoT lo_i = bl[i], lo_j = bl[j];
                    bool i_gt_j = key_gt<LoT>(hi_i, lo_i, hi_j, lo_j);
                    if (desc ? !i_gt_j : i_gt_j) {
                        if (hi_i != hi_j || lo_i != lo_j) {
                            bh[i] = hi_j; bh[j] = hi_i;
                            bl[i] = lo_j; bl[j] = lo_i;
                        }
                    }
                }
                __syncthreads();
            }
        }
    }
};

}  // namespace b200rag
