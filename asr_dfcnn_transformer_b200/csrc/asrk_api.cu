// Version and error strings of the asrk C ABI (include/asrk.h).
#include <atomic>

#include "asrk_common.cuh"

namespace asrk {
static std::atomic<unsigned long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace asrk

extern "C" unsigned long long asrk_launch_count(void) { return asrk::g_launches.load(std::memory_order_relaxed); }

extern "C" int asrk_version(void) { return 100; }   // 0.1.0

extern "C" const char* asrk_error_string(int code) {
    switch (code) {
        case ASRK_OK: return "ok";
        case ASRK_E_BADARG: return "bad argument (null pointer, negative size or unknown enum)";
        case ASRK_E_SHAPE: return "shape not supported by the kernels";
        case ASRK_E_ALIGN: return "pointer or stride not aligned as documented";
        case ASRK_E_WORKSPACE: return "workspace too small or not 256-byte aligned";
        case ASRK_E_CUDA: return "CUDA runtime / kernel launch failure";
        default: return "unknown asrk status";
    }
}
