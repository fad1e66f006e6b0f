// Library plumbing: version, thread-local error string, SM count, TMA descriptor encoding.
#include "common.cuh"
#include "../../include/movedepth_b200.h"

#include <cudaTypedefs.h>
#include <mutex>

namespace mvd {

static thread_local char g_err[512] = "";
static int g_sms = 0;

char* err_buf() { return g_err; }

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(static_cast<int>(e), "%s: %s", what, cudaGetErrorString(e));
    return 0;
}

int sm_count() {
    if (g_sms == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            g_sms = n;
        else
            return 148;
    }
    return g_sms;
}

cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// cuTensorMapEncodeTiled is a driver entry point; fetch it through the runtime so the library
// does not link libcuda directly (it must load, and export its symbols, on a GPU-less host).
static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    });
    return fn;
}

int make_nhwc32_tensor_map(CUtensorMap* map, const float* base, int B, int h, int w, int box_w, int box_h) {
    auto fn = encode_fn();
    if (!fn) return fail(-2, "cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t gdim[4] = {32u, static_cast<cuuint64_t>(w), static_cast<cuuint64_t>(h), static_cast<cuuint64_t>(B)};
    cuuint64_t gstr[3] = {128u, static_cast<cuuint64_t>(w) * 128u, static_cast<cuuint64_t>(w) * h * 128u};
    cuuint32_t box[4] = {32u, static_cast<cuuint32_t>(box_w), static_cast<cuuint32_t>(box_h), 1u};
    cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(-3, "cuTensorMapEncodeTiled failed (CUresult %d)", static_cast<int>(r));
    return 0;
}

// Generic fp32 tiled descriptor (no swizzle): dims/box innermost first, strides in bytes for dims 1..rank-1.
int make_f32_tensor_map(CUtensorMap* map, const float* base, int rank, const uint64_t* dims, const uint64_t* strides,
                        const uint32_t* box, int swizzle_bytes) {
    auto fn = encode_fn();
    if (!fn) return fail(-2, "cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bx[5], estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        estr[i] = 1u;
        if (i > 0) gstr[i - 1] = strides[i - 1];
    }
    const CUtensorMapSwizzle sw = swizzle_bytes == 12832 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                                  : swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                  : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<float*>(base), gdim, gstr, bx, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(-3, "cuTensorMapEncodeTiled failed (CUresult %d)", static_cast<int>(r));
    return 0;
}

}  // namespace mvd

extern "C" {

int mvd_version(void) { return MVD_ABI_VERSION; }
const char* mvd_last_error_string(void) { return mvd::err_buf(); }
int mvd_sm_count(void) { return mvd::sm_count(); }

void* mvd_event_create(void) {
    cudaEvent_t e = nullptr;
    if (cudaEventCreateWithFlags(&e, cudaEventDefault) != cudaSuccess) return nullptr;
    return e;
}

int mvd_event_elapsed_ms(void* start, void* stop, float* ms) {
    MVD_REQUIRE(start && stop && ms, "null argument");
    cudaError_t e = cudaEventElapsedTime(ms, reinterpret_cast<cudaEvent_t>(start), reinterpret_cast<cudaEvent_t>(stop));
    if (e != cudaSuccess) return mvd::fail(static_cast<int>(e), "cudaEventElapsedTime: %s", cudaGetErrorString(e));
    return 0;
}

int mvd_event_destroy(void* event) {
    if (event) cudaEventDestroy(reinterpret_cast<cudaEvent_t>(event));
    return 0;
}

int mvd_event_record(void* event, void* stream, int external) {
    MVD_REQUIRE(event != nullptr, "null event");
    cudaError_t e = cudaEventRecordWithFlags(reinterpret_cast<cudaEvent_t>(event), mvd::as_stream(stream),
                                             external ? cudaEventRecordExternal : cudaEventRecordDefault);
    if (e != cudaSuccess) return mvd::fail(static_cast<int>(e), "cudaEventRecordWithFlags: %s", cudaGetErrorString(e));
    return 0;
}

}
