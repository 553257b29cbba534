// tc_common.cuh -- constants, PTX wrappers and the warp-level candidate compaction shared by the tensor-core scan kernels.
#pragma once
#include <cuda.h>
#include <cstdio>

#include "common.cuh"

namespace b200rag {

// ----------------------------------------------------------------------------------------------- tile configuration
constexpr int TC_BM = 128;             // queries per CTA tile (UMMA M, TMEM lanes)
constexpr int TC_BN = 256;             // corpus rows per accumulator (UMMA N)
constexpr int TC_BK = 64;              // K elements per stage = one 128-byte swizzle atom of 16-bit data
constexpr int TC_STAGES = 4;
constexpr int TC_Q_BYTES = TC_BM * TC_BK * 2;      // 16 KB
constexpr int TC_X_BYTES = TC_BN * TC_BK * 2;      // 32 KB
constexpr int TC_STAGE_BYTES = TC_Q_BYTES + TC_X_BYTES;
constexpr int TC_THREADS = 192;        // warp 0 producer, warp 1 MMA, warps 2..5 epilogue
constexpr int TC_EPI_WARPS = 4;
constexpr int TC_TMEM_COLS = 512;
constexpr int TC_MAX_C = 1280;         // candidate-buffer capacity limit (compaction scratch in smem)
constexpr int TC_FALLBACK_BATCH = 32;
constexpr size_t TC_SMEM_LIMIT = 232448;  // 227 KB of dynamic shared memory per CTA on sm_100
constexpr int TC_QG_SPAN_MAX = 16;     // v3 tile-major order: per-query epilogue state of up to 16 block pairs lives in shared memory

__host__ __device__ inline int tc_kprime(int k) { int s = k / 4 > 28 ? k / 4 : 28; return (k + s + 31) / 32 * 32; }
__host__ __device__ inline int tc_bufcap(int kp) { return 2 * kp; }

// ----------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 6000000000ll) {     // ~3 s at 2 GHz
            printf("b200rag: mbarrier wait timed out (block %d thread %d parity %u)\n", blockIdx.x, threadIdx.x, parity);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "elect.sync _|P1, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}" : "=r"(pred));
    return pred != 0;
}
// Issue a 32-column load without waiting; pair with tmem_ld32_wait(r) before the first use of r.
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}
// Wait for every tcgen05.ld issued so far.  The registers are in/out operands so that the compiler cannot move a use
// of them above the wait.
__device__ __forceinline__ void tmem_ld32_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
        : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
          "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
          "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
          "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
        :: "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor of a K-major, 128-byte-swizzled operand tile whose rows are 128 bytes apart and whose
// 8-row groups are 1024 bytes apart (what TMA SWIZZLE_128B writes for a 64-element-wide 16-bit box).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);       // start address, 16-byte units
    d |= (uint64_t)1 << 16;                            // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                            // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
    return d;
}
// Instruction descriptor: D=f32, A=B=dtype (0 f16 / 1 bf16), both K-major, N=n, M=128.
__host__ __device__ inline uint32_t umma_idesc(int dtype, int n) {
    return (1u << 4) | ((uint32_t)dtype << 7) | ((uint32_t)dtype << 10) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(TC_BM >> 4) << 24);
}


// ----------------------------------------------------------------------------------------------- cluster helpers
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// ----------------------------------------------------------------------------------------------- CTA-pair (cta_group::2) helpers
// Address of `local_smem_addr` in the shared memory of CTA `cta` of this cluster (shared::cluster window).
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(cta));
    return r;
}
// TMA tile load into OUR shared memory whose completion bytes are credited to an mbarrier that may live in the
// peer CTA of the pair (bar_cluster_addr is a shared::cluster address, see mapa_u32).
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster_addr) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
// D[tmem, 256 x N over both CTAs] (+)= A[smem, 128 rows per CTA] * B[smem, N/2 rows per CTA]^T; issued by the leader CTA only.
__device__ __forceinline__ void umma_f16_ss_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// Arrive (once the pair's MMAs issued so far retire) on the barrier at this offset in every CTA of `mask`.
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
// Arrive on a barrier that may live in the peer CTA.  Default semantics (release at CTA scope): what travels through this
// barrier is "the TMEM accumulator is drained", which tcgen05.fence::before_thread_sync orders.  A cluster-scope release here
// made every hand-back wait until the thread's candidate stores to global memory were visible cluster-wide (17 % of the
// scan kernel's stall samples in capture r1g).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Wait on a LOCAL barrier whose arrivals come from other CTAs of the cluster (acquire at cluster scope).
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (clock64() - t0 > 6000000000ll) {
            printf("b200rag: cluster mbarrier wait timed out (block %d thread %d parity %u)\n", blockIdx.x, threadIdx.x, parity);
            __trap();
        }
    }
}
// Instruction descriptor with an explicit M (256 for a CTA pair).
__host__ __device__ inline uint32_t umma_idesc_mn(int dtype, int m, int n) {
    return (1u << 4) | ((uint32_t)dtype << 7) | ((uint32_t)dtype << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

// ----------------------------------------------------------------------------------------------- candidate compaction
// Candidate entry: (fp32 score bits << 32) | local row.  A query's buffer holds up to `cap` entries in global memory; when
// it is nearly full the warp keeps the kp greatest scores and the kp-th greatest becomes the query's new threshold
// (entries equal to the threshold are kept only as far as needed to reach kp).

// Register-resident version: cap = 32*EPL entries, EPL per lane.  ~1.5k cycles instead of ~20k for the shared-memory one.
template <int EPL>
__device__ __forceinline__ float warp_compact_reg(unsigned long long* buf, int n, int kp, int lane) {
    unsigned long long e[EPL];
    uint32_t key[EPL];
#pragma unroll
    for (int t = 0; t < EPL; ++t) {
        const int j = t * 32 + lane;
        e[t] = j < n ? buf[j] : 0ull;
        key[t] = j < n ? mono32(__uint_as_float((uint32_t)(e[t] >> 32))) : 0u;      // 0 sorts below every real key
    }
    uint32_t T = 0;
#pragma unroll 1
    for (int bit = 31; bit >= 0; --bit) {
        const uint32_t cand = T | (1u << bit);
        int c = 0;
#pragma unroll
        for (int t = 0; t < EPL; ++t) c += key[t] >= cand;
        c = __reduce_add_sync(0xffffffffu, c);
        if (c >= kp) T = cand;
    }
    int gt = 0;
#pragma unroll
    for (int t = 0; t < EPL; ++t) gt += key[t] > T;
    gt = __reduce_add_sync(0xffffffffu, gt);
    const int allowed_eq = kp - gt;
    int out = 0, eq_used = 0;
    const unsigned below = (1u << lane) - 1;
#pragma unroll
    for (int t = 0; t < EPL; ++t) {
        const bool is_eq = key[t] == T && T != 0;
        const unsigned m_eq = __ballot_sync(0xffffffffu, is_eq);
        const bool keep = key[t] > T || (is_eq && eq_used + __popc(m_eq & below) < allowed_eq);
        const unsigned m_keep = __ballot_sync(0xffffffffu, keep);
        if (keep) buf[out + __popc(m_keep & below)] = e[t];
        out += __popc(m_keep);
        eq_used += __popc(m_eq);
    }
    __syncwarp();
    return unmono32(T);
}

// Generic version for large buffers: keys staged in a per-warp shared-memory scratch (32-bit shared address).
__device__ __forceinline__ float warp_compact_smem(unsigned long long* buf, int n, int kp, uint32_t scratch_addr, int lane) {
    for (int j = lane; j < n; j += 32) sts_u32(scratch_addr + 4 * j, mono32(__uint_as_float((uint32_t)(buf[j] >> 32))));
    __syncwarp();
    uint32_t T = 0;
    for (int bit = 31; bit >= 0; --bit) {
        const uint32_t cand = T | (1u << bit);
        int c = 0;
        for (int j = lane; j < n; j += 32) c += lds_u32(scratch_addr + 4 * j) >= cand;
        c = __reduce_add_sync(0xffffffffu, c);
        if (c >= kp) T = cand;
    }
    int gt = 0;
    for (int j = lane; j < n; j += 32) gt += lds_u32(scratch_addr + 4 * j) > T;
    gt = __reduce_add_sync(0xffffffffu, gt);
    const int allowed_eq = kp - gt;
    int out = 0, eq_used = 0;
    const unsigned below = (1u << lane) - 1;
    for (int base = 0; base < n; base += 32) {
        const int j = base + lane;
        const bool valid = j < n;
        unsigned long long e = 0;
        uint32_t key = 0;
        if (valid) { e = buf[j]; key = lds_u32(scratch_addr + 4 * j); }
        const bool is_eq = valid && key == T;
        const unsigned m_eq = __ballot_sync(0xffffffffu, is_eq);
        const bool keep = (valid && key > T) || (is_eq && eq_used + __popc(m_eq & below) < allowed_eq);
        const unsigned m_keep = __ballot_sync(0xffffffffu, keep);
        __syncwarp();
        if (keep) buf[out + __popc(m_keep & below)] = e;
        out += __popc(m_keep);
        eq_used += __popc(m_eq);
        __syncwarp();
    }
    return unmono32(T);
}

__device__ __forceinline__ float warp_compact(unsigned long long* buf, int n, int kp, int cap, uint32_t scratch_addr, int lane) {
    switch (cap) {
        case 64: return warp_compact_reg<2>(buf, n, kp, lane);
        case 128: return warp_compact_reg<4>(buf, n, kp, lane);
        case 192: return warp_compact_reg<6>(buf, n, kp, lane);
        case 256: return warp_compact_reg<8>(buf, n, kp, lane);
        case 320: return warp_compact_reg<10>(buf, n, kp, lane);
        case 384: return warp_compact_reg<12>(buf, n, kp, lane);
        case 512: return warp_compact_reg<16>(buf, n, kp, lane);
        default: return warp_compact_smem(buf, n, kp, scratch_addr, lane);
    }
}

// Shared per-query threshold: key 0 means "nobody has k' candidates yet" (-inf).
__device__ __forceinline__ float gthr_load(const unsigned int* p) {
    unsigned int k;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(k) : "l"(p) : "memory");
    return k ? unmono32(k) : -CUDART_INF_F;
}

// Cycle-counter slots written per CTA when ScanParams::stats is set (one elected lane per role).
enum { ST_MMA_TOTAL = 0, ST_MMA_WAIT_FULL, ST_MMA_WAIT_TEMPTY, ST_MMA_WAIT_Q, ST_PROD_TOTAL, ST_PROD_WAIT_EMPTY,
       ST_EPI_TOTAL, ST_EPI_WAIT_TFULL, ST_EPI_COMPACT, ST_EPI_QLOAD, ST_EPI_NCOMPACT, ST_EPI_NSLOW, ST_N = 16 };
#define ST_T0(var) long long var = clock64()
#define ST_ADD(acc, var) acc += clock64() - var

// ----------------------------------------------------------------------------------------------- epilogue building blocks
// One accumulator (n_groups * 32 columns) of the FULL scan: a thread owns one query (TMEM lane), compares 32 scores at a
// time against the query's running threshold and appends the rare survivors to its candidate buffer; when a buffer is
// nearly full the warp compacts it to kp entries and raises the threshold.
struct EpiCounters { long long compact = 0, ncompact = 0, nslow = 0; };

// One group of 32 scores against the query's threshold.
//   VAR 0: a 32-long chain of compares decides "anything above?"; the survivor path then walks all 32 columns.
//   VAR 1: maxima of the four 8-column sub-groups (a tree) decide, and the survivor path only walks the sub-groups whose
//          maximum beats the threshold.  A warp takes the survivor path when ANY of its 32 queries has a survivor in the
//          group, which is most groups on small shards (8 k' survivors spread over few tiles), so its cost matters.
template <int VAR>
__device__ __forceinline__ void epi_filter_group(uint32_t (&r)[32], int c, bool partial, int64_t row0, int64_t n_rows, float& thr,
                                                 int& cnt, unsigned long long* buf, EpiCounters& ec, const uint32_t* row_mask) {
    if (partial) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (row0 + c * 32 + i >= n_rows) r[i] = 0xff800000u;      // -inf: never passes
    }
    if (VAR == 0) {
        bool any = false;
#pragma unroll
        for (int i = 0; i < 32; ++i) any |= __uint_as_float(r[i]) > thr;
        if (any) {
            ++ec.nslow;
            const uint32_t rbase = (uint32_t)(row0 + c * 32);
            // metadata filter: consulted only here, for the rare rows that beat the threshold (the threshold itself is the
            // k'-th best among ALLOWED rows, because only allowed rows are ever appended or sampled)
            const uint32_t allowed = row_mask ? __ldg(row_mask + (rbase >> 5)) : 0xffffffffu;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                if (__uint_as_float(r[i]) > thr && ((allowed >> i) & 1u)) {
                    buf[cnt] = ((unsigned long long)r[i] << 32) | (unsigned long long)(rbase + i);
                    ++cnt;
                }
            }
        }
    } else {
        float m8[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float a0 = fmaxf(__uint_as_float(r[8 * j + 0]), __uint_as_float(r[8 * j + 1]));
            float a1 = fmaxf(__uint_as_float(r[8 * j + 2]), __uint_as_float(r[8 * j + 3]));
            float a2 = fmaxf(__uint_as_float(r[8 * j + 4]), __uint_as_float(r[8 * j + 5]));
            float a3 = fmaxf(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7]));
            m8[j] = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
        }
        if (fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])) > thr) {
            ++ec.nslow;
            const uint32_t rbase = (uint32_t)(row0 + c * 32);
            const uint32_t allowed = row_mask ? __ldg(row_mask + (rbase >> 5)) : 0xffffffffu;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (m8[j] > thr) {
#pragma unroll
                    for (int i = 8 * j; i < 8 * j + 8; ++i) {
                        if (__uint_as_float(r[i]) > thr && ((allowed >> i) & 1u)) {
                            buf[cnt] = ((unsigned long long)r[i] << 32) | (unsigned long long)(rbase + i);
                            ++cnt;
                        }
                    }
                }
            }
        }
    }
}

// make room: a lane appends at most 32 entries per column group
__device__ __forceinline__ void epi_make_room(float& thr, int& cnt, unsigned long long* buf, unsigned int* my_gthr, int kp, int cap,
                                              uint32_t scratch, int lane, EpiCounters& ec) {
    unsigned need = __ballot_sync(0xffffffffu, cnt > cap - 32);
    if (need) {
        ST_T0(tc0);
        ec.ncompact += __popc(need);
        while (need) {
            const int L = __ffs(need) - 1;
            need &= need - 1;
            unsigned long long* b = reinterpret_cast<unsigned long long*>(
                __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(buf), L));
            const int n = __shfl_sync(0xffffffffu, cnt, L);
            const float t = warp_compact(b, n, kp, cap, scratch, lane);
            if (lane == L) {
                cnt = kp;
                thr = fmaxf(thr, t);
                atomicMax(my_gthr, mono32(t));
            }
        }
        ST_ADD(ec.compact, tc0);
    }
}

// The 32-column TMEM loads are software pipelined: while one group of 32 scores is compared, the next is in flight.
template <int NG, int VAR>
__device__ __forceinline__ void epi_filter_tile(uint32_t taddr, int64_t row0, int64_t n_rows, float& thr, int& cnt,
                                                unsigned long long* buf, unsigned int* my_gthr, int kp, int cap,
                                                uint32_t scratch, int lane, EpiCounters& ec, const uint32_t* row_mask) {
    static_assert(NG % 2 == 0, "column groups are processed in pairs");
    const bool partial = row0 + NG * 32 > n_rows;
    uint32_t ra[32], rb[32];
    tmem_ld32_issue(taddr, ra);
#pragma unroll 1
    for (int c = 0; c < NG; c += 2) {
        epi_make_room(thr, cnt, buf, my_gthr, kp, cap, scratch, lane, ec);
        tmem_ld32_wait(ra);
        tmem_ld32_issue(taddr + (c + 1) * 32, rb);
        epi_filter_group<VAR>(ra, c, partial, row0, n_rows, thr, cnt, buf, ec, row_mask);
        epi_make_room(thr, cnt, buf, my_gthr, kp, cap, scratch, lane, ec);
        tmem_ld32_wait(rb);
        if (c + 2 < NG) tmem_ld32_issue(taddr + (c + 2) * 32, ra);
        epi_filter_group<VAR>(rb, c + 1, partial, row0, n_rows, thr, cnt, buf, ec, row_mask);
    }
}

// End of a work item of the full scan: leave at most kp entries per query and publish the threshold.
__device__ __forceinline__ void epi_filter_finish(int& cnt, unsigned long long* buf, unsigned int* my_gthr, int kp, int cap,
                                                  uint32_t scratch, int lane) {
    unsigned need = __ballot_sync(0xffffffffu, cnt > kp);
    while (need) {
        const int L = __ffs(need) - 1;
        need &= need - 1;
        unsigned long long* b = reinterpret_cast<unsigned long long*>(
            __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(buf), L));
        const int n = __shfl_sync(0xffffffffu, cnt, L);
        const float t = warp_compact(b, n, kp, cap, scratch, lane);
        if (lane == L) {
            cnt = kp;
            atomicMax(my_gthr, mono32(t));
        }
    }
}

// SAMPLE pass: no buffers, no compaction.  Each thread keeps, in registers, the TC_SAMPLE_R greatest 32-row GROUP
// MAXIMA it has seen, sorted descending.  The r-th greatest group maximum is a lower bound of the r-th greatest score
// (r distinct rows reach it), and equals it unless two of the top r rows share a group -- so the threshold derived
// from it is never too high because of the grouping, at worst a little low (a few more survivors in the full scan).
constexpr int TC_SAMPLE_R = 16;
constexpr int TC_T0_SAMPLE_MULT = 4;   // tier-0 re-scan: rows above the sampled threshold, in k'

__device__ __forceinline__ void epi_sample_tile(uint32_t taddr, int n_groups, int64_t row0, int64_t n_rows,
                                                float (&top)[TC_SAMPLE_R], const uint32_t* row_mask) {
    const bool partial = row0 + n_groups * 32 > n_rows;
#pragma unroll 1
    for (int c = 0; c < n_groups; ++c) {
        uint32_t r[32];
        tmem_ld32(taddr + c * 32, r);
        if (partial) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (row0 + c * 32 + i >= n_rows) r[i] = 0xff800000u;
        }
        if (row_mask && row0 + c * 32 < n_rows) {               // filtered search: only allowed rows seed the threshold
            const uint32_t allowed = __ldg(row_mask + ((row0 + c * 32) >> 5));
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (!((allowed >> i) & 1u)) r[i] = 0xff800000u;
        }
        float m = __uint_as_float(r[0]);
#pragma unroll
        for (int i = 1; i < 32; ++i) m = fmaxf(m, __uint_as_float(r[i]));
        if (m > top[TC_SAMPLE_R - 1]) {
#pragma unroll
            for (int i = 0; i < TC_SAMPLE_R; ++i) {
                const float hi = fmaxf(top[i], m);
                m = fminf(top[i], m);
                top[i] = hi;
            }
        }
    }
}
__device__ __forceinline__ void epi_sample_finish(const float (&top)[TC_SAMPLE_R], unsigned long long* buf) {
#pragma unroll
    for (int i = 0; i < TC_SAMPLE_R; ++i) buf[i] = (unsigned long long)__float_as_uint(top[i]) << 32;
}

// Scan-kernel parameters shared by both kernels.
struct ScanParams {
    int64_t n_rows;
    int n_q;
    int dim;
    int n_kblocks;       // ceil(dim / 64)
    int n_tiles;         // tiles this launch visits (ceil(n_rows / tile rows) for a full scan)
    int tile_stride;     // visited tile t is corpus tile t * tile_stride (1 = full scan, >1 = strided sample pass)
    int nqb;             // query blocks (padded to a multiple of the cluster size in v2)
    int n_chunks;
    int n_items;
    int qg_span;         // v3: query-block PAIRS one work item covers back to back for every corpus tile (1 = one pair per item)
    int kprime;
    int cap;             // candidate buffer capacity per (chunk, query)
    int sample;          // 1 = strided SAMPLE pass: the epilogue keeps only the TC_SAMPLE_R best 32-row group maxima per query
    uint32_t idesc;
    unsigned long long* cand;   // [n_chunks][nqb][128][cap]  (score bits << 32 | local row)
    int* cand_cnt;              // [n_chunks][nqb][128]
    unsigned int* gthr;         // [nqb*128] shared per-query threshold keys (mono32), monotone via atomicMax
    const uint16_t* queries;    // query block matrix (padded to whole blocks)
    const uint32_t* row_mask;   // optional filter: bit (row & 31) of word (row >> 5) set = row allowed; NULL = no filter
    unsigned long long* stats;  // optional [gridDim.x][16] cycle counters (debug/profiling), may be NULL
    const int* gate;            // optional device scalar: only queries [0, *gate) exist (dense_tc3.cu; NULL = all n_q)
};

}  // namespace b200rag
