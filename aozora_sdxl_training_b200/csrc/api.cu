// C-ABI plumbing shared by every kernel file: thread-local error string, device properties,
// and the cuTensorMapEncodeTiled entry point (resolved through the runtime so that the library
// does not link libcuda directly and can be dlopen'ed on a box without a GPU).
#include "common.cuh"
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

namespace aoz {

static thread_local char g_err[512] = "";
long long g_launch_count = 0;
int g_pdl = 0;       // measured on B200 (r01): 150.0 ms/step with PDL vs 148.7 without under CUDA-graph replay -> off by default

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess) return 148;
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 148;
        n = p.multiProcessorCount;
    }
    return n;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

// bf16 tensor map, 128-byte swizzle, zero OOB fill.  dims/strides innermost first; strides in BYTES for
// dims 1..rank-1 (dim 0 is contiguous).  elem_strides = traversal strides (1 = dense).
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, const uint32_t* elem_strides) {
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled not available (no CUDA driver?)");
        return AOZ_ERR_CUDA;
    }
    cuuint64_t gd[5], gs[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) {
        gd[i] = dims[i];
        bx[i] = box[i];
        es[i] = elem_strides ? elem_strides[i] : 1;
    }
    for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): rank=%d dims=[%llu,%llu,%llu,%llu] stride1=%llu box=[%u,%u,%u,%u] base=%p",
                  (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                  (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
                  (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), box[0], rank > 1 ? box[1] : 0,
                  rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, base);
        return AOZ_ERR_CUDA;
    }
    return AOZ_OK;
}

}  // namespace aoz

extern "C" {

const char* aoz_last_error(void) { return aoz::g_err; }

int aoz_abi_version(void) { return 1; }

int aoz_sm_count(void) { return aoz::sm_count(); }

long long aoz_launch_count(void) { return aoz::g_launch_count; }

// 1: chain kernels with programmatic dependent launch; 0 (default): plain stream order
int aoz_set_pdl(int on) { aoz::g_pdl = on; return AOZ_OK; }

}  // extern "C"
