// Error plumbing and device checks for the C ABI (include/mdir_b200.h).
#include "common.cuh"

namespace mdir {
static thread_local std::string g_err;
unsigned long long g_launches = 0;
int g_scan_max_ctas = 0;
int g_finalize_cluster = 1;
int g_finalize_stage_cap = 0;
void set_error(const std::string& s) { g_err = s; }
int fail_arg(const char* what) {
    g_err = std::string("invalid argument: ") + what;
    return MDIR_E_ARG;
}
}  // namespace mdir

extern "C" int mdir_abi_version(void) { return MDIR_ABI_VERSION; }
extern "C" const char* mdir_last_error(void) { return mdir::g_err.c_str(); }

extern "C" uint64_t mdir_launch_count(void) { return mdir::g_launches; }

extern "C" int mdir_tune(int key, int value) {
    if (key == MDIR_TUNE_SCAN_MAX_CTAS && value >= 0) { mdir::g_scan_max_ctas = value; return 0; }
    if (key == MDIR_TUNE_FINALIZE_CLUSTER && (value == 0 || value == 1)) { mdir::g_finalize_cluster = value; return 0; }
    if (key == MDIR_TUNE_FINALIZE_STAGE_CAP && value >= 0) { mdir::g_finalize_stage_cap = value; return 0; }
    return mdir::fail_arg("mdir_tune(key, value)");
}

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
