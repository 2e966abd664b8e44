// C-ABI bookkeeping for libsn2_b200.so: version, error strings, last CUDA error (thread local).
#include "sn2_common.cuh"
#include <stdio.h>

static thread_local char g_last_err[512] = "";

void sn2_set_cuda_error(cudaError_t e, const char *where)
{
    snprintf(g_last_err, sizeof(g_last_err), "%s: %s (%s)", where, cudaGetErrorString(e), cudaGetErrorName(e));
}

extern "C" int sn2_abi_version(void) { return SN2_ABI_VERSION; }

extern "C" const char *sn2_last_cuda_error(void) { return g_last_err; }

extern "C" const char *sn2_error_string(int code)
{
    switch (code) {
    case SN2_OK: return "ok";
    case SN2_EINVAL: return "invalid argument (null pointer, bad shape or size)";
    case SN2_EUNSUPPORTED: return "unsupported configuration for this build";
    case SN2_ECUDA: return "CUDA error (see sn2_last_cuda_error)";
    default: return "unknown error code";
    }
}
