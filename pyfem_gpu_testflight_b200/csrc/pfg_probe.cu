// Live roofline denominators measured on the device a bench runs on ("also report FP64 FMA-pipe
// utilisation"; appendix B: "no fp64 figure recorded -- measure one").  Diagnostics, not part of the assembly path.
#include "pfg_internal.cuh"

namespace pfg {

// Every thread runs 16 independent DFMA chains: enough instruction-level parallelism to keep the FP64 pipe of an SM
// sub-partition issuing with 8 resident warps.  The result is written so that nothing is optimised away.
__global__ void __launch_bounds__(256) k_probe_dfma(int iters, double seed, double* __restrict__ out) {
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + (double)(threadIdx.x + i);
    const double m = 1.0 - 1e-9, c = 1e-9;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fma(a[i], m, c);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * (size_t)blockDim.x + threadIdx.x] = s;
}

}  // namespace pfg

using namespace pfg;

extern "C" int pfg_probe_fp64(int device, int iters, double* tflops_out, double* ms_out) {
    if (!tflops_out || iters <= 0) {
        set_error("pfg_probe_fp64: invalid argument");
        return PFG_ERR_INVALID;
    }
    PFG_CUDA_TRY(cudaSetDevice(device));
    int sms = 0;
    PFG_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int ctas = sms * 8, threads = 256;
    double* out = nullptr;
    PFG_CUDA_TRY(cudaMalloc(&out, (size_t)ctas * threads * sizeof(double)));
    cudaEvent_t e0, e1;
    PFG_CUDA_TRY(cudaEventCreate(&e0));
    PFG_CUDA_TRY(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {  // first repetition warms the clocks up; best of the rest
        PFG_CUDA_TRY(cudaEventRecord(e0, 0));
        k_probe_dfma<<<ctas, threads>>>(iters, 1.0, out);
        PFG_CUDA_TRY(cudaEventRecord(e1, 0));
        PFG_CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        PFG_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    PFG_CUDA_TRY(cudaGetLastError());
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    const double flops = 2.0 * 64.0 * (double)iters * (double)ctas * threads;  // 64 DFMA per thread and iteration
    *tflops_out = flops / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return PFG_OK;
}
