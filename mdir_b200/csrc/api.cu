// Error plumbing and device checks for the C ABI (include/mdir_b200.h).
#include "common.cuh"

namespace mdir {
static thread_local std::string g_err;
unsigned long long g_launches = 0;
void set_error(const std::string& s) { g_err = s; }
int fail_arg(const char* what) {
    g_err = std::string("invalid argument: ") + what;
    return MDIR_E_ARG;
}
}  // namespace mdir

extern "C" int mdir_abi_version(void) { return MDIR_ABI_VERSION; }
extern "C" const char* mdir_last_error(void) { return mdir::g_err.c_str(); }

extern "C" uint64_t mdir_launch_count(void) { return mdir::g_launches; }

extern "C" int mdir_device_check(void) {
    int dev = 0;
    MDIR_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    MDIR_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) {
        mdir::set_error(std::string("libmdir_b200 is built for sm_100a only; device is ") + prop.name);
        return MDIR_E_DEVICE;
    }
    return 0;
}
