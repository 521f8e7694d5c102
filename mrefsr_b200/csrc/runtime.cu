// mrefsr_b200/csrc/runtime.cu -- error plumbing, device queries, TMA descriptor encoding, device arena.
#include <stdarg.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "../../include/mrefsr_b200.h"

namespace mrefsr {

static thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

void count_launches(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

int sm_count() {   // of the CURRENT device (a process may drive several GPUs)
    static std::atomic<int> cached[64];
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (dev >= 0 && dev < 64 && (n = cached[dev].load(std::memory_order_relaxed)) > 0) return n;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    if (dev >= 0 && dev < 64) cached[dev].store(n, std::memory_order_relaxed);
    return n;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    });
    return fn;
}

int make_tensor_map_3d(CUtensorMap* map, CUtensorMapDataType dtype, int elem_bytes, const void* base, uint64_t d0,
                       uint64_t d1, uint64_t d2, uint32_t box0, uint32_t box1, CUtensorMapSwizzle swizzle) {
    PFN_encodeTiled fn = encode_fn();
    MREFSR_CHECK(fn != nullptr, ERR_NOT_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {d0 * elem_bytes, d0 * d1 * elem_bytes};
    cuuint32_t box[3] = {box0, box1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, dtype, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MREFSR_CHECK(r == CUDA_SUCCESS, ERR_BAD_ARG,
                 "cuTensorMapEncodeTiled failed (%d): dims %llu x %llu x %llu box %u x %u", (int)r,
                 (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, box0, box1);
    return 0;
}

int make_tensor_map_3d_strided(CUtensorMap* map, CUtensorMapDataType dtype, int elem_bytes, const void* base, uint64_t d0,
                               uint64_t d1, uint64_t d2, uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0,
                               uint32_t box1, CUtensorMapSwizzle swizzle) {
    PFN_encodeTiled fn = encode_fn();
    MREFSR_CHECK(fn != nullptr, ERR_NOT_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
    cuuint32_t box[3] = {box0, box1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, dtype, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MREFSR_CHECK(r == CUDA_SUCCESS, ERR_BAD_ARG,
                 "cuTensorMapEncodeTiled failed (%d): dims %llu x %llu x %llu strides %llu / %llu box %u x %u", (int)r,
                 (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
                 (unsigned long long)stride1_bytes, (unsigned long long)stride2_bytes, box0, box1);
    return 0;
}

int make_tensor_map_4d(CUtensorMap* map, CUtensorMapDataType dtype, int elem_bytes, const void* base, uint64_t d0,
                       uint64_t d1, uint64_t d2, uint64_t d3, uint32_t box0, uint32_t box1, uint32_t box2,
                       CUtensorMapSwizzle swizzle) {
    PFN_encodeTiled fn = encode_fn();
    MREFSR_CHECK(fn != nullptr, ERR_NOT_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[4] = {d0, d1, d2, d3};
    cuuint64_t strides[3] = {d0 * elem_bytes, d0 * d1 * elem_bytes, d0 * d1 * d2 * elem_bytes};
    cuuint32_t box[4] = {box0, box1, box2, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, dtype, 4, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MREFSR_CHECK(r == CUDA_SUCCESS, ERR_BAD_ARG,
                 "cuTensorMapEncodeTiled failed (%d): dims %llu x %llu x %llu x %llu box %u x %u x %u", (int)r,
                 (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2, (unsigned long long)d3, box0,
                 box1, box2);
    return 0;
}

// ---------------------------------------------------------------- per-kernel timing
static std::atomic<int> g_timing_on{0};
struct EvPair {
    cudaEvent_t a, b;
    int id;
};
static std::mutex g_timing_mu;
static std::vector<EvPair> g_ev_live;   // recorded, not yet read
static std::vector<EvPair> g_ev_free;   // recycled
static thread_local EvPair g_ev_open[MREFSR_K_COUNT];
static thread_local bool g_ev_is_open[MREFSR_K_COUNT] = {false};

void timing_begin(int id, cudaStream_t st) {
    if (!g_timing_on.load(std::memory_order_relaxed) || id < 0 || id >= MREFSR_K_COUNT) return;
    EvPair p;
    {
        std::lock_guard<std::mutex> lk(g_timing_mu);
        if (!g_ev_free.empty()) {
            p = g_ev_free.back();
            g_ev_free.pop_back();
        } else {
            if (cudaEventCreate(&p.a) != cudaSuccess || cudaEventCreate(&p.b) != cudaSuccess) return;
        }
    }
    p.id = id;
    cudaEventRecord(p.a, st);
    g_ev_open[id] = p;
    g_ev_is_open[id] = true;
}

void timing_end(int id, cudaStream_t st) {
    if (id < 0 || id >= MREFSR_K_COUNT || !g_ev_is_open[id]) return;
    g_ev_is_open[id] = false;
    cudaEventRecord(g_ev_open[id].b, st);
    std::lock_guard<std::mutex> lk(g_timing_mu);
    g_ev_live.push_back(g_ev_open[id]);
}

// ---------------------------------------------------------------- device arena for the *_host entry points
static std::mutex g_arena_mu;
static void* g_arena[64] = {nullptr};       // one arena per device: memory of device 0 is no use to a call on device 1
static size_t g_arena_bytes[64] = {0};

int arena_get(size_t bytes, void** out) {
    int dev = 0;
    MREFSR_CUDA(cudaGetDevice(&dev));
    MREFSR_CHECK(dev >= 0 && dev < 64, ERR_UNSUPPORTED, "arena: device index %d out of range", dev);
    std::lock_guard<std::mutex> lk(g_arena_mu);
    if (bytes > g_arena_bytes[dev]) {
        if (g_arena[dev]) {
            cudaDeviceSynchronize();
            cudaFree(g_arena[dev]);
            g_arena[dev] = nullptr;
            g_arena_bytes[dev] = 0;
        }
        size_t want = align_up(bytes + bytes / 8, (size_t)1 << 21);
        MREFSR_CUDA(cudaMalloc(&g_arena[dev], want));
        g_arena_bytes[dev] = want;
    }
    *out = g_arena[dev];
    return 0;
}

}  // namespace mrefsr

extern "C" {
int mrefsr_abi_version(void) { return MREFSR_ABI_VERSION; }
const char* mrefsr_last_error(void) { return mrefsr::get_error(); }
int mrefsr_sm_count(void) { return mrefsr::sm_count(); }
unsigned long long mrefsr_launch_count(void) { return mrefsr::g_launches.load(); }
void mrefsr_timing_enable(int on) { mrefsr::g_timing_on.store(on ? 1 : 0); }
int mrefsr_timing_read(double* ms_out, unsigned long long* launches_out, int n) {
    using namespace mrefsr;
    if (!ms_out || !launches_out || n < MREFSR_K_COUNT) {
        set_error("timing_read: need arrays of at least %d entries", (int)MREFSR_K_COUNT);
        return ERR_BAD_ARG;
    }
    for (int i = 0; i < n; ++i) {
        ms_out[i] = 0.0;
        launches_out[i] = 0;
    }
    std::lock_guard<std::mutex> lk(g_timing_mu);
    for (auto& p : g_ev_live) {
        float ms = 0.f;
        if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
            ms_out[p.id] += ms;
            launches_out[p.id] += 1;
        }
        g_ev_free.push_back(p);
    }
    g_ev_live.clear();
    return 0;
}
void mrefsr_arena_release(void) {
    std::lock_guard<std::mutex> lk(mrefsr::g_arena_mu);
    int cur = 0;
    cudaGetDevice(&cur);
    for (int dev = 0; dev < 64; ++dev)
        if (mrefsr::g_arena[dev]) {
            cudaSetDevice(dev);
            cudaDeviceSynchronize();
            cudaFree(mrefsr::g_arena[dev]);
            mrefsr::g_arena[dev] = nullptr;
            mrefsr::g_arena_bytes[dev] = 0;
        }
    cudaSetDevice(cur);
}
}
