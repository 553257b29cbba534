// dense_finish.cu -- finish stage of the tensor-core dense scan, second generation  (K2).
//
// Per query: merge the per-chunk candidate lists the scan epilogue wrote to the k' best by tensor-core score, RE-SCORE them in
// the canonical fp64 arithmetic (bit-identical to oracle/exact_scan.c), rank by (score desc, row asc), emit the top k and
// PROVE completeness (see dense_tc.cu's header).  Round 1 ran one 256-thread CTA per query around a block-level streaming
// top-k: 63 M warp instructions for 1024 queries, three quarters of them in the selection's barriers, histograms and
// bitonic passes (profiles/r1: 0.18 ms, 37 % issue utilisation, 3 CTAs per SM).  This generation:
//   * one 128-thread CTA per query, ~10 KB of shared memory -> 16 CTAs per SM, no block-level selection at all;
//   * warp 0 streams the chunk lists through a 512-entry buffer and keeps the k' greatest with the same register-resident
//     radix select the scan epilogue uses (tc_common.cuh: warp_compact) -- typically two or three compactions per query;
//     meanwhile warps 1..3 stage the query as fp32 and sum its squared norm;
//   * one THREAD per candidate row for the re-score: the thread walks its row with 16-byte loads and keeps the eight
//     canonical fp64 lanes in registers.  For fp16 the product of two stored values is exact in fp32, so it is formed with one
//     FMUL and widened once (one conversion per element instead of two); bf16 products can leave the fp32 range and take the
//     fp64 path;
//   * all-pairs rank count in shared memory (k' <= 640 -> at most 3200 compares per thread), no sort.
// Which candidates survive a tie at the k'-th tensor-core score is immaterial: every dropped row has a tensor-core score
// <= m either way, so the proof -- and with it the exact result -- does not depend on it.
#include "finish.cuh"

namespace b200rag {

constexpr int F2_THREADS = 128;

__host__ __device__ inline int finish2_sel_cap(int kprime) { return 2 * kprime <= 512 ? 512 : 2 * kprime; }

size_t finish2_smem_bytes(int dim, int kprime) {
    const int sel = finish2_sel_cap(kprime);
    return (size_t)dim * 4 + (size_t)sel * 8 + (size_t)sel * 4 + (size_t)kprime * (8 + 4) + 64;
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
    const int sel_cap = finish2_sel_cap(p.kprime);
    float* qf = reinterpret_cast<float*>(smem);                                           // [dim] the query, exact in fp32
    unsigned long long* buf = reinterpret_cast<unsigned long long*>(qf + p.dim);          // [sel_cap] (score bits << 32) | row
    uint32_t* scratch = reinterpret_cast<uint32_t*>(buf + sel_cap);                        // [sel_cap] keys of the smem select
    double* exact = reinterpret_cast<double*>(scratch + sel_cap);                          // [kprime]
    uint32_t* rows = reinterpret_cast<uint32_t*>(exact + p.kprime);                        // [kprime]
    __shared__ int s_n;
    __shared__ float s_m;
    __shared__ double s_q2[F2_THREADS / 32];
    __shared__ float s_err;
    __shared__ double s_ek_sh;

    if (warp == 0) {
        // ---- 1. the k' best by tensor-core score over all chunks of this query
        int cnt = 0;
        float thr = -CUDART_INF_F;
        bool compacted = false;
        for (int c = 0; c < p.n_chunks; ++c) {
            const size_t slot = ((size_t)(c * p.nqb + qb)) * TC_BM + ql;
            const int n_c = __ldg(p.cand_cnt + slot);
            const unsigned long long* src = p.cand + slot * p.cap;
            for (int base = 0; base < n_c; base += 32) {
                const int i = base + lane;
                unsigned long long e = 0ull;
                bool pass = false;
                if (i < n_c) {
                    e = __ldg(src + i);
                    pass = !compacted || __uint_as_float((uint32_t)(e >> 32)) > thr;
                }
                const unsigned bal = __ballot_sync(0xffffffffu, pass);
                if (pass) buf[cnt + __popc(bal & ((1u << lane) - 1u))] = e;
                cnt += __popc(bal);
                if (cnt > sel_cap - 32) {
                    __syncwarp();
                    thr = fmaxf(thr, warp_compact(buf, cnt, p.kprime, sel_cap, smem_u32(scratch), lane));
                    cnt = p.kprime;
                    compacted = true;
                }
            }
        }
        __syncwarp();
        if (cnt > p.kprime) {
            thr = fmaxf(thr, warp_compact(buf, cnt, p.kprime, sel_cap, smem_u32(scratch), lane));
            cnt = p.kprime;
            compacted = true;
        }
        if (lane == 0) {
            s_n = cnt;
            // m: every row outside the candidate set has tensor-core score <= m (-inf if nothing was ever dropped): the scan
            // drops a row only below a threshold it pushed to gthr, the merge above only at or below its last threshold
            const unsigned int gk = p.gthr[slot_q];
            s_m = fmaxf(compacted ? thr : -CUDART_INF_F, gk ? unmono32(gk) : -CUDART_INF_F);
            s_err = 0.f;
            s_ek_sh = -CUDART_INF;
        }
    } else {
        // ---- meanwhile: the query as fp32 (16-bit values are exact in fp32) and its squared norm
        double q2 = 0.0;
        for (int d = tid - 32; d < p.dim; d += F2_THREADS - 32) {
            const double v = bits_to_double<DTYPE>(p.queries[(size_t)q * p.dim + d]);
            qf[d] = (float)v;
            q2 = fma(v, v, q2);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) q2 += __shfl_xor_sync(0xffffffffu, q2, off);
        if (lane == 0) s_q2[warp] = q2;
    }
    __syncthreads();
    const int n = s_n;

    // ---- 2. exact canonical re-score, one thread per candidate row
    float my_err = 0.f;
    const int n_vec = p.dim >> 3;
    const float4* q4 = reinterpret_cast<const float4*>(qf);
    for (int i = tid; i < n; i += F2_THREADS) {
        const unsigned long long e = buf[i];
        const uint32_t row = (uint32_t)e;
        const uint4* x = reinterpret_cast<const uint4*>(p.corpus + (size_t)row * p.dim);
        double p0 = 0, p1 = 0, p2 = 0, p3 = 0, p4 = 0, p5 = 0, p6 = 0, p7 = 0;
#pragma unroll 4
        for (int c = 0; c < n_vec; ++c) {
            const uint4 v = __ldg(x + c);
            const float4 qa = q4[2 * c], qb4 = q4[2 * c + 1];
            if (DTYPE == B200RAG_F16) {
                const float2 x0 = __half22float2(*reinterpret_cast<const __half2*>(&v.x));
                const float2 x1 = __half22float2(*reinterpret_cast<const __half2*>(&v.y));
                const float2 x2 = __half22float2(*reinterpret_cast<const __half2*>(&v.z));
                const float2 x3 = __half22float2(*reinterpret_cast<const __half2*>(&v.w));
                // fp16 x fp16 is exact in fp32 (22 significant bits, exponents within range): one rounding-free FMUL, one
                // widening, one fp64 add == fma(q, x, p) of the canonical definition, bit for bit
                p0 = __dadd_rn(p0, (double)__fmul_rn(qa.x, x0.x)); p1 = __dadd_rn(p1, (double)__fmul_rn(qa.y, x0.y));
                p2 = __dadd_rn(p2, (double)__fmul_rn(qa.z, x1.x)); p3 = __dadd_rn(p3, (double)__fmul_rn(qa.w, x1.y));
                p4 = __dadd_rn(p4, (double)__fmul_rn(qb4.x, x2.x)); p5 = __dadd_rn(p5, (double)__fmul_rn(qb4.y, x2.y));
                p6 = __dadd_rn(p6, (double)__fmul_rn(qb4.z, x3.x)); p7 = __dadd_rn(p7, (double)__fmul_rn(qb4.w, x3.y));
            } else {
                double a, b;
                unpack2<DTYPE>(v.x, a, b); p0 = fma((double)qa.x, a, p0); p1 = fma((double)qa.y, b, p1);
                unpack2<DTYPE>(v.y, a, b); p2 = fma((double)qa.z, a, p2); p3 = fma((double)qa.w, b, p3);
                unpack2<DTYPE>(v.z, a, b); p4 = fma((double)qb4.x, a, p4); p5 = fma((double)qb4.y, b, p5);
                unpack2<DTYPE>(v.w, a, b); p6 = fma((double)qb4.z, a, p6); p7 = fma((double)qb4.w, b, p7);
            }
        }
        const double t = __dadd_rn(__dadd_rn(__dadd_rn(p0, p1), __dadd_rn(p2, p3)), __dadd_rn(__dadd_rn(p4, p5), __dadd_rn(p6, p7)));
        exact[i] = t;
        rows[i] = row;
        my_err = fmaxf(my_err, fabsf((float)((double)__uint_as_float((uint32_t)(e >> 32)) - t)));
    }
    if (p.err_max) atomicMax(reinterpret_cast<int*>(&s_err), __float_as_int(my_err));       // non-negative floats order as ints
    __syncthreads();

    // ---- 3. rank by (exact desc, row asc) and emit: all-pairs rank count
    const int kk = min(p.k, n);
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
    for (int i = kk + tid; i < p.k; i += F2_THREADS) {
        p.out_scores[(size_t)q * p.k + i] = -CUDART_INF;
        p.out_ids[(size_t)q * p.k + i] = -1;
    }
    __syncthreads();
    if (tid == 0) {
        const double q2 = s_q2[1] + s_q2[2] + s_q2[3];
        const double eps = 2.0 * (double)p.dim * 1.1920928955078125e-07 * sqrt(q2) * p.row_norm_bound;
        const float m = s_m;
        // proven complete iff nothing was dropped (m = -inf) or the k-th exact score clears m + eps
        const bool proven = (m == -CUDART_INF_F) || (n >= p.k && s_ek_sh > (double)m + eps);
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
