// NVTX ranges around the C-ABI entry points (tracing): `nsys`/`ncu --nvtx` attribute kernels to the call
// that launched them (b200s_chol_factorize, b200s_klu_refactor_batch, ...).  Header-only NVTX v3: no library to link, and
// a no-op when no profiler is attached.
#pragma once
#include <nvtx3/nvToolsExt.h>

namespace b200s {
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};
}  // namespace b200s
#define B200S_NVTX(name) ::b200s::NvtxRange nvtx_range_(name)
