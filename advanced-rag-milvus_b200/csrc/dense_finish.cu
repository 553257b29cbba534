// dense_finish.cu -- finish stage of the tensor-core dense scan, second generation  (K2).
//
// Per query: merge the per-chunk candidate lists the scan epilogue wrote to the k' best by tensor-core score, RE-SCORE them in
// the canonical fp64 arithmetic (bit-identical to oracle/exact_scan.c), rank by (score desc, row asc), emit the top k and
// PROVE completeness (see dense_tc.cu's header).  Round 1 ran one 256-thread CTA per query around a block-level streaming
// top-k: 63 M warp instructions for 1024 queries, three quarters of them in the selection's barriers, histograms and
// bitonic passes (profiles/r1: 0.18 ms, 37 % issue utilisation, 3 CTAs per SM).  This generation:
//   * one 128-thread CTA per query, ~24 KB of shared memory -> 8 CTAs per SM: 1024 queries finish in one wave;
//   * all threads gather the query's survivors (~8 k', a handful per chunk) into a working set that aliases the row staging
//     region; whenever it holds more than k' entries the whole block selects the k' best with a radix select over
//     range-normalised keys (block_select below).  (The first version of this generation let warp 0 stream the survivors 32 at a
//     time through a register-resident radix select: 59 us of selection per batch against 37 us now, and 1 ms at k' = 640.)
//   * re-score from rows staged 16 at a time, in 384-element segments, with warp-wide 16-byte cp.async copies (coalesced
//     requests), eight threads per row, one per canonical fp64 lane.  (One thread walking a whole row straight from global
//     memory -- 16-byte requests from 32 different rows per warp instruction -- measured 262 us, slower than round 1.)
//   * rank: all-pairs count in shared memory up to 256 candidates, bitonic sort of (score, row) pairs above.
// Split mode (launch_finish3, finish_version 3) runs the same selection, then the re-score as its own grid and a rank kernel.
// Which candidates survive a tie at the k'-th tensor-core score is immaterial: every dropped row has a tensor-core score
// <= m either way, so the proof -- and with it the exact result -- does not depend on it.
#include "finish.cuh"

namespace b200rag {

constexpr int F2_THREADS = 128;
constexpr int F2_ROWS = F2_THREADS / 8;     // candidate rows re-scored per batch: eight threads share a row
constexpr int F2_GATHER = 1024;             // survivors gathered per round (8 KB, aliased with the row staging buffer)

constexpr int F2_SEG = 384;                 // elements of a row staged at a time (768 bytes): the staging buffer does not grow with dim,
                                            // so eight CTAs fit an SM and 1024 queries finish in ONE wave (at 16 whole 768-d rows:
                                            // five per SM, 1.4 waves)
__host__ __device__ inline int finish2_seg(int dim) { return dim < F2_SEG ? dim : F2_SEG; }
// The staging region doubles as the selection's working set: k' running winners + one round of gathered survivors.
__host__ __device__ inline size_t finish2_stage_bytes(int dim, int kprime) {
    size_t stage = (size_t)F2_ROWS * ((size_t)finish2_seg(dim) * 2 + 16);
    const size_t work = ((size_t)kprime + F2_GATHER / 2) * 8;
    if (work > stage) stage = work;
    if (kprime > 256) {                                  // the bitonic rank sorts (score, row) pairs of the next power of two here
        size_t np2 = 512;
        while (np2 < (size_t)kprime) np2 <<= 1;
        if (np2 * 12 > stage) stage = np2 * 12;
    }
    return stage;
}

size_t finish2_smem_bytes(int dim, int kprime) {
    return (size_t)dim * 8 + (size_t)kprime * (8 + 8 + 4) + 256 * 4 + 256 * 4 + finish2_stage_bytes(dim, kprime) + 64;
}

// Block-wide selection of the kprime best of work[0, cnt) by tensor-core score (the high 32 bits of an entry, compared as
// order-preserving keys).  The scores of one query share their exponent, so the keys are first made relative to the set's
// minimum and shifted up to fill 32 bits: the four 8-bit radix passes that narrow the key of the kprime-th best entry then
// see evenly filled histograms and plain shared-memory atomics do not pile up (a first version aggregated them per warp with
// match.any -- a third of the kernel's stall samples).  Warp 0 picks the digit after each pass; one more pass moves the winners
// to out[0, kprime): everything above the threshold key and as many of its ties as are still needed (which ties does not
// matter: the caller's bound m covers every dropped entry, and a result is only returned if the proof holds).  Returns the
// threshold score.  All F2_THREADS threads call; cnt > kprime.
__device__ float block_select(const unsigned long long* work, int cnt, int kprime, unsigned long long* out, int* hist, int* s_sel) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t kmin = 0xffffffffu, kmax = 0u;
    for (int i = tid; i < cnt; i += F2_THREADS) {
        const uint32_t key = mono32(__uint_as_float((uint32_t)(work[i] >> 32)));
        kmin = min(kmin, key);
        kmax = max(kmax, key);
    }
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    if (lane == 0) { hist[warp] = (int)kmin; hist[8 + warp] = (int)kmax; }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < F2_THREADS / 32; ++w) { kmin = min(kmin, (uint32_t)hist[w]); kmax = max(kmax, (uint32_t)hist[8 + w]); }
    __syncthreads();
    const int lz = kmax > kmin ? __clz(kmax - kmin) : 0;           // (all keys equal: every digit is 0, the tie rule takes kprime of them)
    uint32_t prefix = 0u;
    int need = kprime;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        for (int b = tid; b < 256; b += F2_THREADS) hist[b] = 0;
        __syncthreads();
        for (int i = tid; i < cnt; i += F2_THREADS) {
            const uint32_t key = (mono32(__uint_as_float((uint32_t)(work[i] >> 32))) - kmin) << lz;
            if (pass == 0 || (key >> (shift + 8)) == (prefix >> (shift + 8))) atomicAdd(&hist[(key >> shift) & 255u], 1);
        }
        __syncthreads();
        if (warp == 0) {
            // the digit d of the kprime-th best: the largest d with  sum_{b >= d} hist[b] >= need
            int h[8], mine = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { h[j] = hist[lane * 8 + j]; mine += h[j]; }
            int above = mine;                                    // inclusive suffix sum over lanes (lane 31 holds the top digits)
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_down_sync(0xffffffffu, above, o);
                if (lane + o < 32) above += u;
            }
            const int higher = above - mine;                     // entries in the lanes above this one
            if (higher < need && above >= need) {                // exactly one lane
                int acc = higher, d = 7;
                for (; d > 0; --d) {
                    if (acc + h[d] >= need) break;
                    acc += h[d];
                }
                s_sel[0] = lane * 8 + d;
                s_sel[1] = need - acc;                           // still to take among the entries that carry this digit
            }
        }
        __syncthreads();
        prefix |= (uint32_t)s_sel[0] << shift;
        need = s_sel[1];
        __syncthreads();
    }
    if (tid == 0) { s_sel[2] = 0; s_sel[3] = 0; }
    __syncthreads();
    const int n_above = kprime - need;
    for (int i0 = 0; i0 < cnt; i0 += F2_THREADS) {
        const int i = i0 + tid;
        unsigned long long e = 0ull;
        uint32_t key = 0u;
        if (i < cnt) { e = work[i]; key = (mono32(__uint_as_float((uint32_t)(e >> 32))) - kmin) << lz; }
        const bool gt = i < cnt && key > prefix, eq = i < cnt && key == prefix;
        const unsigned bg = __ballot_sync(0xffffffffu, gt), be = __ballot_sync(0xffffffffu, eq);
        int base_g = 0, base_e = 0;
        if (lane == 0) {
            if (bg) base_g = atomicAdd(&s_sel[2], __popc(bg));
            if (be) base_e = atomicAdd(&s_sel[3], __popc(be));
        }
        base_g = __shfl_sync(0xffffffffu, base_g, 0);
        base_e = __shfl_sync(0xffffffffu, base_e, 0);
        if (gt) out[base_g + __popc(bg & ((1u << lane) - 1u))] = e;
        if (eq) {
            const int t = base_e + __popc(be & ((1u << lane) - 1u));
            if (t < need) out[n_above + t] = e;
        }
    }
    __syncthreads();
    return unmono32(kmin + (prefix >> lz));
}

template <int DTYPE>
__global__ void __launch_bounds__(F2_THREADS) dense_finish2_kernel(const FinishParams p) {
    extern __shared__ __align__(16) char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // slot = position of the query in THIS launch's query block matrix (candidate buffers, thresholds); q = the query the
    // results belong to.  They differ in the tier-0 re-scan, which packs the flagged queries of the first pass.
    const int slot_q = blockIdx.x;
    if (p.gate && slot_q >= __ldg(p.gate)) return;
    const int q = p.q_list ? __ldg(p.q_list + slot_q) : slot_q;
    const int qb = slot_q / TC_BM, ql = slot_q % TC_BM;
    double* qd = reinterpret_cast<double*>(smem);                                         // [dim] the query in fp64
    unsigned long long* buf = reinterpret_cast<unsigned long long*>(qd + p.dim);          // [kprime] the selection: (score bits << 32) | row
    double* exact = reinterpret_cast<double*>(buf + p.kprime);                             // [kprime]
    uint32_t* rows = reinterpret_cast<uint32_t*>(exact + p.kprime);                        // [kprime]
    int* s_cnt = reinterpret_cast<int*>(rows + p.kprime);                                  // [<= 256] survivors per chunk
    int* hist = s_cnt + 256;                                                               // [256] radix histogram of the selection
    char* stage = reinterpret_cast<char*>(hist + 256);
    stage = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(stage) + 15) & ~uintptr_t(15));
    __shared__ int s_sel[4], s_wcnt;
    __shared__ int s_n, s_total;
    __shared__ float s_m;
    __shared__ double s_tau;                             // bound on the tensor score of candidates dropped before the re-score
    __shared__ double s_q2[F2_THREADS / 32];
    __shared__ float s_err;
    __shared__ double s_ek_sh;

    // how many survivors did the scan leave in every chunk of this query?  (all threads, one round trip)
    for (int c = tid; c < p.n_chunks; c += F2_THREADS) s_cnt[c] = __ldg(p.cand_cnt + ((size_t)(c * p.nqb + qb)) * TC_BM + ql);
    __syncthreads();
    if (warp == 0) {                                     // exclusive prefix of the counts, in place (<= 256 chunks)
        int carry = 0;
        for (int c0 = 0; c0 < p.n_chunks; c0 += 32) {
            const int c = c0 + lane;
            const int v = c < p.n_chunks ? s_cnt[c] : 0;
            int incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += u;
            }
            if (c < p.n_chunks) s_cnt[c] = carry + incl - v;
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) s_total = carry;
    }
    __syncthreads();
    const int total = s_total;
    // ---- 1. the k' best by tensor-core score over all chunks of this query.  The survivors -- a handful per chunk, ~8 k' in
    //         all -- are copied into the working set in shared memory (the region the re-score later stages rows in) with
    //         independent loads, one round trip per round; whenever the set holds more than k' entries the whole block selects
    //         the k' best (block_select) and later rounds only admit entries above the threshold that left.
    unsigned long long* work = reinterpret_cast<unsigned long long*>(stage);
    const int round = (int)(finish2_stage_bytes(p.dim, p.kprime) / 8) - p.kprime;      // survivors admitted per round
    float thr = -CUDART_INF_F;
    bool compacted = false;
    if (tid == 0) s_wcnt = 0;
    __syncthreads();
    for (int g0 = 0; g0 < total; g0 += round) {
        const int g_n = min(round, total - g0);
        for (int i0 = 0; i0 < g_n; i0 += F2_THREADS) {
            const int i = i0 + tid;
            unsigned long long e = 0ull;
            bool pass = false;
            if (i < g_n) {
                const int pos = g0 + i;
                int lo = 0, hi = p.n_chunks - 1;         // last chunk whose prefix is <= pos
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (s_cnt[mid] <= pos) lo = mid;
                    else hi = mid - 1;
                }
                e = __ldg(p.cand + (((size_t)(lo * p.nqb + qb)) * TC_BM + ql) * p.cap + (pos - s_cnt[lo]));
                pass = !compacted || __uint_as_float((uint32_t)(e >> 32)) > thr;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, pass);
            int base = 0;
            if (lane == 0 && bal) base = atomicAdd(&s_wcnt, __popc(bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (pass) work[base + __popc(bal & ((1u << lane) - 1u))] = e;
        }
        __syncthreads();
        const int cnt = s_wcnt;
        if (cnt > p.kprime) {
            thr = fmaxf(thr, block_select(work, cnt, p.kprime, buf, hist, s_sel));
            compacted = true;
            for (int i = tid; i < p.kprime; i += F2_THREADS) work[i] = buf[i];
            if (tid == 0) s_wcnt = p.kprime;
            __syncthreads();
        }
    }
    {
        const int cnt = s_wcnt;
        if (!compacted) for (int i = tid; i < cnt; i += F2_THREADS) buf[i] = work[i];
        if (tid == 0) {
            s_n = cnt;
            // m: every row outside the candidate set has tensor-core score <= m (-inf if nothing was ever dropped): the scan
            // drops a row only below a threshold it pushed to gthr, the selection above only at or below its last threshold
            const unsigned int gk = p.gthr[slot_q];
            s_m = fmaxf(compacted ? thr : -CUDART_INF_F, gk ? unmono32(gk) : -CUDART_INF_F);
            s_err = 0.f;
            s_ek_sh = -CUDART_INF;
            s_tau = -CUDART_INF;
        }
    }
    if (!p.fin_sel) {
        // ---- the query in fp64 and its squared norm
        double q2 = 0.0;
        for (int d = tid; d < p.dim; d += F2_THREADS) {
            const double v = bits_to_double<DTYPE>(p.queries[(size_t)q * p.dim + d]);
            qd[d] = v;
            q2 = fma(v, v, q2);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) q2 += __shfl_xor_sync(0xffffffffu, q2, off);
        if (lane == 0) s_q2[warp] = q2;
    }
    __syncthreads();
    int n = s_n;
    // ---- 1b. deep candidate sets (the tier-0 re-scan keeps k' = 640 for k = 100): only the candidates whose tensor-core score
    //          is within 3 eps of the k-th best one can reach the exact top k -- a candidate below  tau = t_k - 3 eps  has an
    //          exact score < tau + eps = t_k - 2 eps, while the k best by tensor-core score all have exact scores >= t_k - eps.
    //          The others are dropped before the re-score (the FP64 pipe bounds it: one conversion and one DFMA per element) and
    //          the proof treats them like rows outside the candidate set: bound tau instead of m.
    if (!p.approx && !p.fin_sel && n >= 2 * p.k) {
        const double q2b = s_q2[0] + s_q2[1] + s_q2[2] + s_q2[3];
        const double epsb = 2.0 * (double)p.dim * 1.1920928955078125e-07 * sqrt(q2b) * p.row_norm_bound;
        const float t_k = block_select(buf, n, p.k, reinterpret_cast<unsigned long long*>(exact), hist, s_sel);
        const double tau = (double)t_k - 3.0 * epsb;
        if (tid == 0) s_wcnt = 0;
        __syncthreads();
        for (int i0 = 0; i0 < n; i0 += F2_THREADS) {
            const int i = i0 + tid;
            unsigned long long e = 0ull;
            bool keep = false;
            if (i < n) {
                e = buf[i];
                keep = (double)__uint_as_float((uint32_t)(e >> 32)) >= tau;
            }
            const unsigned bal = __ballot_sync(0xffffffffu, keep);
            int base = 0;
            if (lane == 0 && bal) base = atomicAdd(&s_wcnt, __popc(bal));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (keep) work[base + __popc(bal & ((1u << lane) - 1u))] = e;
        }
        __syncthreads();
        const int kept = s_wcnt;
        if (kept < n) {
            for (int i = tid; i < kept; i += F2_THREADS) buf[i] = work[i];
            if (tid == 0) { s_n = kept; s_tau = tau; }
            n = kept;
        }
        __syncthreads();
    }
    if (p.fin_sel) {                                     // split finish: hand the selection over and leave
        for (int i = tid; i < n; i += F2_THREADS) p.fin_sel[(size_t)slot_q * p.kprime + i] = buf[i];
        if (tid == 0) {
            p.fin_n[slot_q] = n;
            p.fin_m[slot_q] = s_m;
            if (p.err_max) p.err_max[q] = 0.f;
        }
        return;
    }

    // ---- 2. exact canonical re-score.  Rows are staged through shared memory F2_ROWS at a time with 16-byte cp.async copies
    //         (a warp copies whole rows: coalesced 512-byte requests, nothing passes through registers), then eight threads
    //         share a row, one per canonical lane (lane j sums d = j mod 8 in increasing d, exactly as oracle/exact_scan.c
    //         does), and the lanes are combined in the canonical tree with shuffles.
    float my_err = 0.f;
    if (p.approx) {
        // approximate mode: the candidates' tensor-core scores ARE the result scores
        for (int i = tid; i < n; i += F2_THREADS) {
            const unsigned long long e = buf[i];
            rows[i] = (uint32_t)e;
            exact[i] = (double)__uint_as_float((uint32_t)(e >> 32));
        }
    } else {
        constexpr int RB = F2_ROWS;
        const int l8 = tid & 7, grp = tid >> 3;
        const int seg = finish2_seg(p.dim);
        const int stride = seg * 2 + 16;                          // staged row pitch: +16 bytes keeps the rows of a warp on different banks
        const double* qj = qd + l8;
        for (int i0 = 0; i0 < n; i0 += RB) {
            const int rows_here = min(RB, n - i0);
            const int i = i0 + grp;
            const bool valid = i < n;
            const uint16_t* x = reinterpret_cast<const uint16_t*>(stage + (size_t)grp * stride) + l8;
            double acc = 0.0;
            for (int d0 = 0; d0 < p.dim; d0 += seg) {             // (a lane's chain runs on across the segments, in increasing d)
                const int dl = min(seg, p.dim - d0);
                for (int r = warp; r < rows_here; r += F2_THREADS / 32) {
                    const uint32_t row = (uint32_t)buf[i0 + r];
                    const uint4* src = reinterpret_cast<const uint4*>(p.corpus + (size_t)row * p.dim + d0);
                    const uint32_t dst = smem_u32(stage + (size_t)r * stride);
                    for (int c = lane; c < dl / 8; c += 32)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + c * 16), "l"(src + c) : "memory");
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 0;" ::: "memory");
                __syncthreads();
                if (valid) {
#pragma unroll 4
                    for (int d = 0; d < dl; d += 8) acc = fma(qj[d0 + d], bits_to_double<DTYPE>(x[d]), acc);
                }
                __syncthreads();
            }
            double t = __dadd_rn(acc, __shfl_down_sync(0xffffffffu, acc, 1));      // lanes 0,2,4,6: p0+p1, p2+p3, ...
            t = __dadd_rn(t, __shfl_down_sync(0xffffffffu, t, 2));                 // lanes 0,4: (p0+p1)+(p2+p3), ...
            t = __dadd_rn(t, __shfl_down_sync(0xffffffffu, t, 4));                 // lane 0: the canonical score
            if (valid && l8 == 0) {
                const unsigned long long e = buf[i];
                rows[i] = (uint32_t)e;
                exact[i] = t;
                my_err = fmaxf(my_err, fabsf((float)((double)__uint_as_float((uint32_t)(e >> 32)) - t)));
            }
        }
    }
    if (p.err_max) atomicMax(reinterpret_cast<int*>(&s_err), __float_as_int(my_err));       // non-negative floats order as ints
    __syncthreads();

    // ---- 3. rank by (exact desc, row asc) and emit.  Up to 256 candidates: all-pairs rank count (n^2 / 128 comparisons per
    //         thread, no barriers).  Deeper candidate sets (k' = 640 in the tier-0 re-scan: 3200 fp64 comparisons per thread)
    //         are sorted instead: bitonic network over (score, row) pairs in the staging region, which is free again.
    const int kk = min(p.k, n);
    if (n <= 256) {
        for (int i = tid; i < n; i += F2_THREADS) {
            const double e = exact[i];
            const uint32_t r = rows[i];
            int rk = 0;
            for (int j = 0; j < n; ++j) rk += (exact[j] > e) || (exact[j] == e && rows[j] < r);
            if (rk < p.k) {
                p.out_scores[(size_t)q * p.k + rk] = e;
                p.out_ids[(size_t)q * p.k + rk] = p.id_offset + (int64_t)r;
            }
            if (rk == kk - 1) s_ek_sh = e;
        }
    } else {
        int np2 = 512;
        while (np2 < n) np2 <<= 1;
        double* sk = reinterpret_cast<double*>(stage);                     // [np2] scores, padded with -inf
        uint32_t* sr = reinterpret_cast<uint32_t*>(sk + np2);              // [np2] rows, padded with the largest id
        for (int i = tid; i < np2; i += F2_THREADS) {
            sk[i] = i < n ? exact[i] : -CUDART_INF;
            sr[i] = i < n ? rows[i] : 0xffffffffu;
        }
        for (int size = 2; size <= np2; size <<= 1) {
            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                __syncthreads();
                for (int t = tid; t < np2 / 2; t += F2_THREADS) {
                    const int i = 2 * t - (t & (stride - 1));             // lower index of the pair
                    const int j = i + stride;
                    const bool desc = (i & size) == 0;                     // this block of the network sorts best-first
                    const double a = sk[i], b = sk[j];
                    const uint32_t ra = sr[i], rb = sr[j];
                    const bool a_first = (a > b) || (a == b && ra < rb);   // a ranks before b
                    if (a_first != desc) { sk[i] = b; sk[j] = a; sr[i] = rb; sr[j] = ra; }
                }
            }
        }
        __syncthreads();
        for (int i = tid; i < kk; i += F2_THREADS) {
            p.out_scores[(size_t)q * p.k + i] = sk[i];
            p.out_ids[(size_t)q * p.k + i] = p.id_offset + (int64_t)sr[i];
        }
        if (tid == 0 && kk > 0) s_ek_sh = sk[kk - 1];
    }
    for (int i = kk + tid; i < p.k; i += F2_THREADS) {
        p.out_scores[(size_t)q * p.k + i] = -CUDART_INF;
        p.out_ids[(size_t)q * p.k + i] = -1;
    }
    __syncthreads();
    if (tid == 0) {
        const double q2 = s_q2[0] + s_q2[1] + s_q2[2] + s_q2[3];
        const double eps = 2.0 * (double)p.dim * 1.1920928955078125e-07 * sqrt(q2) * p.row_norm_bound;
        const double m = fmax((double)s_m, s_tau);
        // proven complete iff nothing was dropped (m = -inf) or the k-th exact score clears m + eps
        const bool proven = p.approx || (m == -CUDART_INF) || (n >= p.k && s_ek_sh > m + eps);
        const int flag = proven ? 0 : 1;
        if (p.out_flags) p.out_flags[q] = flag | (p.q_list ? 2 : 0);      // bit 1: the query went through the tier-0 re-scan
        if (flag) p.flag_list[atomicAdd(p.n_flagged, 1)] = q;
        if (p.err_max) p.err_max[q] = s_err;
    }
}

int launch_finish2(const FinishParams& fp, int dtype, cudaStream_t st) {
    const size_t smem = finish2_smem_bytes(fp.dim, fp.kprime);
    if (dtype == B200RAG_F16) {
        B200_CUDA_CHECK(cudaFuncSetAttribute(dense_finish2_kernel<B200RAG_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dense_finish2_kernel<B200RAG_F16><<<fp.n_q, F2_THREADS, smem, st>>>(fp); count_launch();
    } else {
        B200_CUDA_CHECK(cudaFuncSetAttribute(dense_finish2_kernel<B200RAG_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dense_finish2_kernel<B200RAG_BF16><<<fp.n_q, F2_THREADS, smem, st>>>(fp); count_launch();
    }
    B200_CUDA_CHECK(cudaGetLastError());
    return B200RAG_OK;
}

// ----------------------------------------------------------------------------------------------- split finish
// finish3 = dense_finish2_kernel in split mode (selection only) + the two kernels below.  The monolithic kernel re-scores a
// query's k' candidate rows 16 at a time inside the query's CTA: k'/16 dependent gather round trips per query, and with 256
// queries x k' = 640 only 256 CTAs to hide them.  Here the re-score is its own grid over (16-row batch, query): every
// candidate row of the whole batch is in flight at once (B = 1024, k' = 128: 8192 CTAs; B = 256, k' = 640: 10240).
template <int DTYPE>
__global__ void __launch_bounds__(F2_THREADS) finish3_rescore_kernel(const FinishParams p) {
    extern __shared__ __align__(16) char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slot_q = blockIdx.y, i0 = blockIdx.x * F2_ROWS;
    const int n = p.fin_n[slot_q];
    if (i0 >= n) return;
    double* qd = reinterpret_cast<double*>(smem);                                         // [dim] the query in fp64
    char* stage = reinterpret_cast<char*>(qd + p.dim);
    const unsigned long long* sel = p.fin_sel + (size_t)slot_q * p.kprime;
    for (int d = tid; d < p.dim; d += F2_THREADS) qd[d] = bits_to_double<DTYPE>(p.queries[(size_t)slot_q * p.dim + d]);
    const int rows_here = min(F2_ROWS, n - i0);
    const int l8 = tid & 7, grp = tid >> 3;
    const int seg = finish2_seg(p.dim);
    const int stride = seg * 2 + 16;
    const double* qj = qd + l8;
    const int i = i0 + grp;
    const bool valid = i < n;
    const uint16_t* x = reinterpret_cast<const uint16_t*>(stage + (size_t)grp * stride) + l8;
    double acc = 0.0;
    for (int d0 = 0; d0 < p.dim; d0 += seg) {                     // (a lane's chain runs on across the segments, in increasing d)
        const int dl = min(seg, p.dim - d0);
        for (int r = warp; r < rows_here; r += F2_THREADS / 32) {
            const uint32_t row = (uint32_t)sel[i0 + r];
            const uint4* src = reinterpret_cast<const uint4*>(p.corpus + (size_t)row * p.dim + d0);
            const uint32_t dst = smem_u32(stage + (size_t)r * stride);
            for (int c = lane; c < dl / 8; c += 32)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + c * 16), "l"(src + c) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                                          // (also publishes qd on the first round)
        if (valid) {
#pragma unroll 4
            for (int d = 0; d < dl; d += 8) acc = fma(qj[d0 + d], bits_to_double<DTYPE>(x[d]), acc);
        }
        __syncthreads();
    }
    double t = __dadd_rn(acc, __shfl_down_sync(0xffffffffu, acc, 1));
    t = __dadd_rn(t, __shfl_down_sync(0xffffffffu, t, 2));
    t = __dadd_rn(t, __shfl_down_sync(0xffffffffu, t, 4));
    if (valid && l8 == 0) {
        p.fin_exact[(size_t)slot_q * p.kprime + i] = t;
        if (p.err_max) {
            const float err = fabsf((float)((double)__uint_as_float((uint32_t)(sel[i] >> 32)) - t));
            atomicMax(reinterpret_cast<int*>(p.err_max + slot_q), __float_as_int(err));       // non-negative floats order as ints
        }
    }
}

constexpr int F3_RANK_THREADS = 256;

template <int DTYPE>
__global__ void __launch_bounds__(F3_RANK_THREADS) finish3_rank_kernel(const FinishParams p) {
    extern __shared__ __align__(16) char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = blockIdx.x;
    double* exact = reinterpret_cast<double*>(smem);                                      // [kprime]
    uint32_t* rows = reinterpret_cast<uint32_t*>(exact + p.kprime);                       // [kprime]
    __shared__ double s_q2[F3_RANK_THREADS / 32];
    __shared__ double s_ek_sh;
    const int n = p.fin_n[q];
    for (int i = tid; i < n; i += F3_RANK_THREADS) {
        exact[i] = p.fin_exact[(size_t)q * p.kprime + i];
        rows[i] = (uint32_t)p.fin_sel[(size_t)q * p.kprime + i];
    }
    double q2 = 0.0;
    for (int d = tid; d < p.dim; d += F3_RANK_THREADS) {
        const double v = bits_to_double<DTYPE>(p.queries[(size_t)q * p.dim + d]);
        q2 = fma(v, v, q2);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) q2 += __shfl_xor_sync(0xffffffffu, q2, off);
    if (lane == 0) s_q2[warp] = q2;
    if (tid == 0) s_ek_sh = -CUDART_INF;
    __syncthreads();
    const int kk = min(p.k, n);
    for (int i = tid; i < n; i += F3_RANK_THREADS) {
        const double e = exact[i];
        const uint32_t r = rows[i];
        int rk = 0;
        for (int j = 0; j < n; ++j) rk += (exact[j] > e) || (exact[j] == e && rows[j] < r);
        if (rk < p.k) {
            p.out_scores[(size_t)q * p.k + rk] = e;
            p.out_ids[(size_t)q * p.k + rk] = p.id_offset + (int64_t)r;
        }
        if (rk == kk - 1) s_ek_sh = e;
    }
    for (int i = kk + tid; i < p.k; i += F3_RANK_THREADS) {
        p.out_scores[(size_t)q * p.k + i] = -CUDART_INF;
        p.out_ids[(size_t)q * p.k + i] = -1;
    }
    __syncthreads();
    if (tid == 0) {
        double q2s = 0.0;                                  // (the proof's epsilon only needs |q| to a few ulps: any summation order)
        for (int w = 0; w < F3_RANK_THREADS / 32; ++w) q2s += s_q2[w];
        const double eps = 2.0 * (double)p.dim * 1.1920928955078125e-07 * sqrt(q2s) * p.row_norm_bound;
        const float m = p.fin_m[q];
        const bool proven = (m == -CUDART_INF_F) || (n >= p.k && s_ek_sh > (double)m + eps);
        const int flag = proven ? 0 : 1;
        if (p.out_flags) p.out_flags[q] = flag;
        if (flag) p.flag_list[atomicAdd(p.n_flagged, 1)] = q;
    }
}

size_t finish3_workspace_bytes(int n_q, int kprime) {
    return align_up((size_t)n_q * kprime * 8, 256) * 2 + align_up((size_t)n_q * 4, 256) * 2;
}

int launch_finish3(FinishParams fp, int dtype, void* fin_ws, cudaStream_t st) {
    char* w = static_cast<char*>(fin_ws);
    const size_t plane = align_up((size_t)fp.n_q * fp.kprime * 8, 256), small = align_up((size_t)fp.n_q * 4, 256);
    fp.fin_sel = reinterpret_cast<unsigned long long*>(w);
    fp.fin_exact = reinterpret_cast<double*>(w + plane);
    fp.fin_n = reinterpret_cast<int*>(w + 2 * plane);
    fp.fin_m = reinterpret_cast<float*>(w + 2 * plane + small);
    int rc = launch_finish2(fp, dtype, st);                       // split mode: selection only
    if (rc) return rc;
    const size_t smem_b = (size_t)fp.dim * 8 + (size_t)F2_ROWS * ((size_t)finish2_seg(fp.dim) * 2 + 16) + 16;
    const size_t smem_c = (size_t)fp.kprime * 12 + 16;
    const dim3 grid_b((fp.kprime + F2_ROWS - 1) / F2_ROWS, fp.n_q);
    if (dtype == B200RAG_F16) {
        B200_CUDA_CHECK(cudaFuncSetAttribute(finish3_rescore_kernel<B200RAG_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
        finish3_rescore_kernel<B200RAG_F16><<<grid_b, F2_THREADS, smem_b, st>>>(fp); count_launch();
        finish3_rank_kernel<B200RAG_F16><<<fp.n_q, F3_RANK_THREADS, smem_c, st>>>(fp); count_launch();
    } else {
        B200_CUDA_CHECK(cudaFuncSetAttribute(finish3_rescore_kernel<B200RAG_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
        finish3_rescore_kernel<B200RAG_BF16><<<grid_b, F2_THREADS, smem_b, st>>>(fp); count_launch();
        finish3_rank_kernel<B200RAG_BF16><<<fp.n_q, F3_RANK_THREADS, smem_c, st>>>(fp); count_launch();
    }
    B200_CUDA_CHECK(cudaGetLastError());
    return B200RAG_OK;
}

// ----------------------------------------------------------------------------------------------- sample threshold
// After the strided SAMPLE pass every (chunk, query) slot holds TC_SAMPLE_R group maxima; the rank-th greatest of a query's
// n_chunks * TC_SAMPLE_R values becomes its starting threshold (dense_tc.cu explains the sizing).  One WARP per query: the
// values are staged in shared memory as order-preserving keys and the rank-th greatest is found with a 32-step bitwise
// search (count of keys >= candidate, one warp reduction per bit).  Round 1 ran a 256-thread CTA with the block-level
// streaming top-k for this.
constexpr int ST2_WARPS = 4;

__global__ void __launch_bounds__(ST2_WARPS * 32)
sample_threshold2_kernel(const unsigned long long* __restrict__ cand, int cap, int nqb, int n_chunks, int rank, int n_q_pad,
                         unsigned int* __restrict__ gthr, const int* __restrict__ gate) {
    extern __shared__ __align__(16) char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = blockIdx.x * ST2_WARPS + warp;
    if (q >= n_q_pad) return;
    if (gate && q >= 2 * TC_BM * ((__ldg(gate) + 2 * TC_BM - 1) / (2 * TC_BM))) return;     // tier 0: pairs of query blocks that exist
    const int qb = q / TC_BM, ql = q % TC_BM;
    const int total = n_chunks * TC_SAMPLE_R;
    uint32_t* keys = reinterpret_cast<uint32_t*>(smem) + (size_t)warp * total;
    for (int e = lane; e < total; e += 32) {
        const int chunk = e / TC_SAMPLE_R, j = e % TC_SAMPLE_R;
        const unsigned long long v = __ldg(cand + (((size_t)(chunk * nqb + qb)) * TC_BM + ql) * cap + j);
        keys[e] = mono32(__uint_as_float((uint32_t)(v >> 32)));
    }
    __syncwarp();
    uint32_t T = 0;
    for (int bit = 31; bit >= 0; --bit) {
        const uint32_t c0 = T | (1u << bit);
        int c = 0;
        for (int e = lane; e < total; e += 32) c += keys[e] >= c0;
        c = __reduce_add_sync(0xffffffffu, c);
        if (c >= rank) T = c0;
    }
    if (lane == 0) gthr[q] = total >= rank ? T : 0u;
}

int launch_sample_threshold2(const unsigned long long* cand, int cap, int nqb, int n_chunks, int rank, unsigned int* gthr,
                             cudaStream_t st, const int* gate) {
    const int n_q_pad = nqb * TC_BM;
    const size_t smem = (size_t)ST2_WARPS * n_chunks * TC_SAMPLE_R * 4;
    B200_CUDA_CHECK(cudaFuncSetAttribute(sample_threshold2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sample_threshold2_kernel<<<(n_q_pad + ST2_WARPS - 1) / ST2_WARPS, ST2_WARPS * 32, smem, st>>>(cand, cap, nqb, n_chunks, rank, n_q_pad, gthr, gate);
    count_launch();
    B200_CUDA_CHECK(cudaGetLastError());
    return B200RAG_OK;
}

}  // namespace b200rag
