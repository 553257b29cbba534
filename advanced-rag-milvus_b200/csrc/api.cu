// api.cu -- error plumbing, device info and row preparation for libb200rag.so.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "common.cuh"

namespace b200rag {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
unsigned long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

// ---- options -----------------------------------------------------------------------------------
static std::atomic<int> g_opt[OPT_COUNT];
static const char* const kOptNames[OPT_COUNT] = {"scan_version", "qg_span", "epi", "no_sample", "stage_rows", "sample_mult",
                                                 "mmr_path", "sparse_slices", "sparse_flags", "no_tier0", "finish_version"};
static struct OptInit { OptInit() { for (auto& o : g_opt) o.store(-1, std::memory_order_relaxed); } } g_opt_init;
int option(Option o, int dflt) {
    const int v = g_opt[o].load(std::memory_order_relaxed);
    return v < 0 ? dflt : v;
}
static int option_index(const char* name) {
    if (!name) return -1;
    for (int i = 0; i < OPT_COUNT; ++i)
        if (strcmp(name, kOptNames[i]) == 0) return i;
    return -1;
}

// ---- caller-owned debug buffers (thread local: two threads profiling at once do not see each other) ----
static thread_local unsigned long long* g_stats_ptr[STATS_KINDS] = {nullptr, nullptr};
static thread_local size_t g_stats_cap[STATS_KINDS] = {0, 0};
unsigned long long* stats_buffer(StatsKind k, size_t need_slots) {
    return (g_stats_ptr[k] && g_stats_cap[k] >= need_slots) ? g_stats_ptr[k] : nullptr;
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return B200RAG_E_CUDA;
}

// ------------------------------------------------------------------------------------------------
// Row preparation: fp32 -> 16-bit, optionally L2-normalised with the canonical arithmetic
//   n2 = sum_d (double)x[d]^2 sequentially in d;  y[d] = rn16( (double)x[d] / sqrt(n2) ).
// One warp per row would break the sequential order, so a thread owns a row; the block stages a tile of rows
// through shared memory so that global reads and writes stay coalesced.
// ------------------------------------------------------------------------------------------------
template <int DTYPE>
__device__ __forceinline__ uint16_t double_to_bits(double v) {
    if (DTYPE == B200RAG_F16) return __half_as_ushort(__double2half(v));
    return __bfloat16_as_ushort(__double2bfloat16(v));
}

constexpr int PREP_ROWS = 8;    // rows per block (one thread each for the sequential fp64 norm; few rows per block so that a
                                // 1024-query batch spreads over 128 SMs instead of 32)
constexpr int PREP_THREADS = 256;

template <int DTYPE>
__global__ void __launch_bounds__(PREP_THREADS) prepare_rows_kernel(const float* __restrict__ in, uint16_t* __restrict__ out,
                                                                    int64_t n_rows, int dim, int normalize) {
    extern __shared__ float s_tile[];            // [PREP_ROWS][dim + 1]
    __shared__ double s_norm[PREP_ROWS];
    const int pitch = dim + 1;
    const int64_t row0 = (int64_t)blockIdx.x * PREP_ROWS;
    const int nr = (int)min((int64_t)PREP_ROWS, n_rows - row0);
    const float* src = in + row0 * dim;
    for (int i = threadIdx.x; i < nr * dim; i += PREP_THREADS) s_tile[(i / dim) * pitch + (i % dim)] = src[i];
    __syncthreads();
    if (threadIdx.x < nr) {
        double n2 = 0.0;
        const float* x = s_tile + threadIdx.x * pitch;
        if (normalize) {
            for (int d = 0; d < dim; ++d) {
                double v = (double)x[d];
                n2 = __dadd_rn(n2, __dmul_rn(v, v));
            }
            s_norm[threadIdx.x] = sqrt(n2);
        } else {
            s_norm[threadIdx.x] = 1.0;
        }
    }
    __syncthreads();
    uint16_t* dst = out + row0 * dim;
    for (int i = threadIdx.x; i < nr * dim; i += PREP_THREADS) {
        int r = i / dim, d = i % dim;
        double v = (double)s_tile[r * pitch + d];
        double nrm = s_norm[r];
        uint16_t bits;
        if (normalize) bits = nrm > 0.0 ? double_to_bits<DTYPE>(__ddiv_rn(v, nrm)) : (uint16_t)0;
        else bits = double_to_bits<DTYPE>(v);
        dst[i] = bits;
    }
}

}  // namespace b200rag

using namespace b200rag;

extern "C" {

const char* b200rag_last_error(void) { return g_err; }

int b200rag_abi_version(void) { return B200RAG_ABI_VERSION; }
uint64_t b200rag_kernel_launch_count(void) { return b200rag::launch_count(); }

int b200rag_set_option(const char* name, int64_t value) {
    const int i = option_index(name);
    B200_REQUIRE(i >= 0, "set_option: unknown option '%s'", name ? name : "(null)");
    g_opt[i].store(value < 0 ? -1 : (int)value, std::memory_order_relaxed);
    return B200RAG_OK;
}

int64_t b200rag_get_option(const char* name) {
    const int i = option_index(name);
    return i < 0 ? -2 : (int64_t)g_opt[i].load(std::memory_order_relaxed);
}

int b200rag_debug_set_stats_buffer(int32_t kind, uint64_t* device_buf, size_t n_slots) {
    B200_REQUIRE(kind >= 0 && kind < STATS_KINDS, "debug_set_stats_buffer: bad kind %d", kind);
    g_stats_ptr[kind] = reinterpret_cast<unsigned long long*>(device_buf);
    g_stats_cap[kind] = device_buf ? n_slots : 0;
    return B200RAG_OK;
}

int b200rag_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    B200_CUDA_CHECK(cudaGetDevice(&dev));
    cudaDeviceProp p;
    B200_CUDA_CHECK(cudaGetDeviceProperties(&p, dev));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    return B200RAG_OK;
}

int b200rag_prepare_rows(const float* in_f32, void* out16, int64_t n_rows, int32_t dim, int32_t dtype,
                         int32_t normalize, void* stream) {
    B200_REQUIRE(n_rows >= 0 && dim > 0 && dim <= 8192, "prepare_rows: bad shape n_rows=%lld dim=%d", (long long)n_rows, dim);
    B200_REQUIRE(dtype == B200RAG_F16 || dtype == B200RAG_BF16, "prepare_rows: bad dtype %d", dtype);
    if (n_rows == 0) return B200RAG_OK;
    B200_REQUIRE(in_f32 && out16, "prepare_rows: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    size_t smem = (size_t)PREP_ROWS * (dim + 1) * sizeof(float);
    int64_t blocks = (n_rows + PREP_ROWS - 1) / PREP_ROWS;
    B200_REQUIRE(blocks < (int64_t)1 << 31, "prepare_rows: too many rows for one call");
    if (dtype == B200RAG_F16) {
        B200_CUDA_CHECK(cudaFuncSetAttribute(prepare_rows_kernel<B200RAG_F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        prepare_rows_kernel<B200RAG_F16><<<(unsigned)blocks, PREP_THREADS, smem, st>>>(in_f32, (uint16_t*)out16, n_rows, dim, normalize); count_launch();
    } else {
        B200_CUDA_CHECK(cudaFuncSetAttribute(prepare_rows_kernel<B200RAG_BF16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        prepare_rows_kernel<B200RAG_BF16><<<(unsigned)blocks, PREP_THREADS, smem, st>>>(in_f32, (uint16_t*)out16, n_rows, dim, normalize); count_launch();
    }
    B200_CUDA_CHECK(cudaGetLastError());
    return B200RAG_OK;
}

}  // extern "C"
