// dense_tc.cu -- placeholder until the tcgen05 scan lands (next commit).
#include "common.cuh"
namespace b200rag {
size_t tensor_workspace_bytes(int64_t, int, int, int) { return 256; }
int run_tensor(const void*, int64_t, int, int, const void*, int, int, int64_t, double*, int64_t*, int32_t*, int, void*, size_t,
               cudaStream_t) {
    set_error("dense_topk: tensor-core path not built");
    return B200RAG_E_UNSUPPORTED;
}
}  // namespace b200rag
