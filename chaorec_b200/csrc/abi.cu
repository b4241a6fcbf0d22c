// ABI bookkeeping + roofline probes.
#include "common.cuh"

namespace chaorec {
namespace {

// 8x8 register outer products with operands held in registers: the same FFMA mix as the hot
// loop of rank_topk.cu without any memory traffic -> the measured fp32 FMA ceiling.
constexpr int kProbeThreads = 256;
__global__ void __launch_bounds__(kProbeThreads, 2) fp32_fma_probe_kernel(int64_t iters, float *sink) {
    float acc[8][8];
    float a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = 1.0f + 1e-7f * static_cast<float>(threadIdx.x + i);
        b[i] = 1.0f - 1e-7f * static_cast<float>(threadIdx.x + 2 * i);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
    }
    for (int64_t it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) s += acc[i][j];
    if (s == 123.456f) sink[0] = s;  // never true; keeps the loop alive
}

}  // namespace
}  // namespace chaorec

using namespace chaorec;

extern "C" int chaorec_abi_version(void) { return CHAOREC_ABI_VERSION; }

extern "C" const char *chaorec_error_string(int code) {
    switch (code) {
        case CHAOREC_OK: return "ok";
        case CHAOREC_E_BADARG: return "bad argument (null pointer, negative or inconsistent size)";
        case CHAOREC_E_UNSUPPORTED: return "unsupported shape (d % 4 != 0, d > 512 for SpMM, K or k outside 1..64, K > I)";
        case CHAOREC_E_WORKSPACE: return "workspace too small";
        case CHAOREC_E_LAUNCH: return "CUDA launch or runtime call failed";
        case CHAOREC_E_ALIGN: return "pointer not aligned (16 bytes for tables, 256 bytes for workspace)";
        default: return "unknown error code";
    }
}

extern "C" int chaorec_probe_fp32_fma(int64_t iters, float *sink, double *flops_out, void *stream) {
    if (!sink || iters < 0) return CHAOREC_E_BADARG;
    const int grid = sm_count() * 2;
    fp32_fma_probe_kernel<<<grid, kProbeThreads, 0, static_cast<cudaStream_t>(stream)>>>(iters, sink);
    if (flops_out) *flops_out = 2.0 * 256.0 * static_cast<double>(iters) * kProbeThreads * grid;
    return launch_status();
}
