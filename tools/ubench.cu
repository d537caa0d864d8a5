// tools/ubench.cu -- developer probe: issue rate / latency of FFMA vs FFMA2 (fma.rn.f32x2) on sm_100a, and of the
// MUFU / LDS / SHFL instructions the step kernels lean on.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

#define REP 4096

template <int ILP>
__global__ void k_ffma(float *out, float a, float b) {
    float x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 0.001f + i;
#pragma unroll 1
    for (int r = 0; r < REP; r++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < ILP; i++) x[i] = fmaf(x[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_ffma2(float *out, float a, float b) {
    float2 x[ILP];
    const float2 A = make_float2(a, a * 1.0001f), B = make_float2(b, b * 0.999f);
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
#pragma unroll 1
    for (int r = 0; r < REP; r++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < ILP; i++) x[i] = __ffma2_rn(x[i], A, B);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// scalar-broadcast form: d.xy = s * b.xy + c.xy
template <int ILP>
__global__ void k_ffma2b(float *out, float a, float b) {
    float2 x[ILP];
    const float2 B = make_float2(b, b * 0.999f);
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
#pragma unroll 1
    for (int r = 0; r < REP; r++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < ILP; i++) x[i] = __ffma2_rn(make_float2(a, a), x[i], B);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_mufu(float *out, float a) {
    float x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 0.001f + i + 1.f;
#pragma unroll 1
    for (int r = 0; r < REP; r++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < ILP; i++) x[i] = rsqrtf(x[i]) + a;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_shfl(float *out) {
    float x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 0.001f + i + 1.f;
#pragma unroll 1
    for (int r = 0; r < REP; r++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < ILP; i++) x[i] += __shfl_xor_sync(0xffffffffu, x[i], 1);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_lds128(float *out) {
    __shared__ float4 sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    float4 x[ILP];
    int idx = threadIdx.x;
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = make_float4(0, 0, 0, 0);
#pragma unroll 1
    for (int r = 0; r < REP; r++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                float4 v = sm[(idx + 128 * i + 7 * u) & 1023];
                x[i].x += v.x; x[i].y += v.y; x[i].z += v.z; x[i].w += v.w;
            }
        idx = (idx + (int)x[0].x) & 1023;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i].x + x[i].y + x[i].z + x[i].w;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_it(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int main() {
    float *out; cudaMalloc(&out, 1 << 24);
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("SMs %d, clock attr %d kHz\n", sms, clk_khz);
    // per-SMSP warp counts: block of 128 threads = 1 warp per scheduler; vary blocks per SM via block size
    for (int wps : {1, 2, 4, 8}) {
        const int threads = 128 * wps > 1024 ? 1024 : 128 * wps, blocks = sms * (128 * wps / threads);
        const double n_inst = (double)REP * 8;   // per ILP slot per warp
#define RUN(NAME, K, ILP, FL)                                                                                     \
        {                                                                                                         \
            float ms = time_it([&] { K<ILP><<<blocks, threads>>> FL; });                                          \
            double cyc = ms * 1e-3 * 1.965e9;                                                                     \
            printf("%-10s ILP %d warps/sched %d: %.3f ms  -> %.2f cyc per warp-instr per scheduler (at 1.965 GHz)\n", NAME, ILP, wps, ms, \
                   cyc / (n_inst * ILP * wps));                                                                   \
        }
        RUN("ffma", k_ffma, 1, (out, 1.0001f, 0.5f)); RUN("ffma", k_ffma, 4, (out, 1.0001f, 0.5f)); RUN("ffma", k_ffma, 8, (out, 1.0001f, 0.5f));
        RUN("ffma2", k_ffma2, 1, (out, 1.0001f, 0.5f)); RUN("ffma2", k_ffma2, 4, (out, 1.0001f, 0.5f)); RUN("ffma2", k_ffma2, 8, (out, 1.0001f, 0.5f));
        RUN("ffma2b", k_ffma2b, 1, (out, 1.0001f, 0.5f)); RUN("ffma2b", k_ffma2b, 4, (out, 1.0001f, 0.5f));
        RUN("mufu", k_mufu, 1, (out, 0.5f)); RUN("mufu", k_mufu, 4, (out, 0.5f));
        RUN("shfl", k_shfl, 1, (out)); RUN("shfl", k_shfl, 4, (out));
        RUN("lds128", k_lds128, 1, (out)); RUN("lds128", k_lds128, 4, (out));
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("done: %s\n", cudaGetErrorString(e));
    return 0;
}
