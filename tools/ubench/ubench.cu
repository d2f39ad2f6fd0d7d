// Micro-benchmarks that size the mh_sweep kernel design on B200 (sm_100a):
// FFMA / FFMA2 issue rates, broadcast LDS.128 rate, mixed LDS+FFMA2, MUFU, DFMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)

__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
    unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
    unsigned long long aa = *reinterpret_cast<const unsigned long long*>(&a);
    unsigned long long bb = *reinterpret_cast<const unsigned long long*>(&b);
    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
    d = *reinterpret_cast<float2*>(&dd);
}

constexpr int ITERS = 2048;

__global__ void k_ffma(float* out, float x) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = threadIdx.x * 1e-3f + i;
    float m = x, c = x * 0.5f;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) a[i] = fmaf(a[i], m, c);
    }
    float s = 0; for (int i = 0; i < 16; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, float x) {
    float2 a[16];
#pragma unroll
    for (int i = 0; i < 16; i++) a[i] = make_float2(threadIdx.x * 1e-3f + i, i);
    float2 m = make_float2(x, x * 1.1f);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) ffma2(a[i], m, a[(i + 1) & 15]);
    }
    float s = 0; for (int i = 0; i < 16; i++) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// broadcast LDS.128 only
__global__ void k_lds(float* out, int stride) {
    __shared__ float4 sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    float4 acc = make_float4(0, 0, 0, 0);
    int base = 0;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            float4 v = sm[(base + i) & 1023];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        base += stride;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}
// mixed: one broadcast LDS.128 feeding NF ffma2 per item, K items
template <int K>
__global__ void k_mix(float* out, int stride) {
    __shared__ float4 sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = make_float4(1e-3f * i, 1e-3f, 2e-3f, 3e-3f);
    __syncthreads();
    float2 acc[K][4];
    float2 e[K][4];
#pragma unroll
    for (int k = 0; k < K; k++)
#pragma unroll
        for (int j = 0; j < 4; j++) { acc[k][j] = make_float2(0, 0); e[k][j] = make_float2(threadIdx.x * 1e-3f + k, j); }
    int base = 0;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            float4 v = sm[(base + i) & 1023];
#pragma unroll
            for (int k = 0; k < K; k++) {
                ffma2(acc[k][(i & 1) * 2], make_float2(v.x, v.y), e[k][i & 3]);
                ffma2(acc[k][(i & 1) * 2 + 1], make_float2(v.z, v.w), e[k][(i + 1) & 3]);
            }
        }
        base += stride;
    }
    float s = 0;
    for (int k = 0; k < K; k++) for (int j = 0; j < 4; j++) s += acc[k][j].x + acc[k][j].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mufu(float* out, float x) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3f + i * 0.1f;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = exp2f(a[i] * x) ;  // fast-math off: use intrinsic below
    }
    float s = 0; for (int i = 0; i < 8; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mufu_raw(float* out, float x) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3f + i * 0.1f;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
    }
    float s = 0; for (int i = 0; i < 8; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// MUFU co-issued with FFMA: 1 mufu per NF ffma
template <int NF>
__global__ void k_mufu_ffma(float* out, float x) {
    float a[8]; float b[NF];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3f + i * 0.1f;
#pragma unroll
    for (int i = 0; i < NF; i++) b[i] = i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
#pragma unroll
            for (int j = 0; j < NF; j++) b[j] = fmaf(b[j], x, 0.5f);
        }
    }
    float s = 0; for (int i = 0; i < 8; i++) s += a[i]; for (int i = 0; i < NF; i++) s += b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dfma(double* out, double x) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) a[i] = fma(a[i], x, 0.5);
    }
    double s = 0; for (int i = 0; i < 8; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount; int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("device %s sms %d clock %d kHz\n", p.name, sms, clk_khz);
    float* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
    const int T = 256;
    for (int wps : {1, 2, 4}) {  // CTAs per SM (each 256 thr = 8 warps)
        int grid = sms * wps;
        double warps = (double)grid * T / 32;
        float ms;
        ms = timeit([&] { k_ffma<<<grid, T>>>(out, 0.999f); });
        printf("[ctas/sm %d] FFMA   : %.3f ms  -> %.2f warp-instr/clk/SM @1.9GHz-equiv (%.1f TFLOP/s)\n", wps, ms, warps * ITERS * 16 / (ms * 1e-3) / sms / 1.9e9, warps * 32 * ITERS * 16 * 2 / (ms * 1e-3) / 1e12);
        ms = timeit([&] { k_ffma2<<<grid, T>>>(out, 0.999f); });
        printf("[ctas/sm %d] FFMA2  : %.3f ms  -> %.2f warp-instr/clk/SM (%.1f TFLOP/s)\n", wps, ms, warps * ITERS * 16 / (ms * 1e-3) / sms / 1.9e9, warps * 32 * ITERS * 16 * 4 / (ms * 1e-3) / 1e12);
        ms = timeit([&] { k_lds<<<grid, T>>>(out, 16); });
        printf("[ctas/sm %d] LDS.128 bcast (+4 FADD): %.3f ms -> %.3f LDS.128/clk/SM\n", wps, ms, warps * ITERS * 16 / (ms * 1e-3) / sms / 1.9e9);
        ms = timeit([&] { k_mix<1><<<grid, T>>>(out, 16); });
        printf("[ctas/sm %d] MIX K=1 (1 LDS.128 : 2 FFMA2): %.3f ms -> %.3f LDS/clk/SM, %.2f FFMA2/clk/SM\n", wps, ms, warps * ITERS * 16 / (ms * 1e-3) / sms / 1.9e9, warps * ITERS * 32 / (ms * 1e-3) / sms / 1.9e9);
        ms = timeit([&] { k_mix<2><<<grid, T>>>(out, 16); });
        printf("[ctas/sm %d] MIX K=2 (1 LDS.128 : 4 FFMA2): %.3f ms -> %.3f LDS/clk/SM, %.2f FFMA2/clk/SM\n", wps, ms, warps * ITERS * 16 / (ms * 1e-3) / sms / 1.9e9, warps * ITERS * 64 / (ms * 1e-3) / sms / 1.9e9);
        ms = timeit([&] { k_mix<3><<<grid, T>>>(out, 16); });
        printf("[ctas/sm %d] MIX K=3 (1 LDS.128 : 6 FFMA2): %.3f ms -> %.3f LDS/clk/SM, %.2f FFMA2/clk/SM\n", wps, ms, warps * ITERS * 16 / (ms * 1e-3) / sms / 1.9e9, warps * ITERS * 96 / (ms * 1e-3) / sms / 1.9e9);
        ms = timeit([&] { k_mix<4><<<grid, T>>>(out, 16); });
        printf("[ctas/sm %d] MIX K=4 (1 LDS.128 : 8 FFMA2): %.3f ms -> %.3f LDS/clk/SM, %.2f FFMA2/clk/SM\n", wps, ms, warps * ITERS * 16 / (ms * 1e-3) / sms / 1.9e9, warps * ITERS * 128 / (ms * 1e-3) / sms / 1.9e9);
        ms = timeit([&] { k_mufu_raw<<<grid, T>>>(out, 0.5f); });
        printf("[ctas/sm %d] MUFU.EX2: %.3f ms -> %.3f warp-instr/clk/SM\n", wps, ms, warps * ITERS * 8 / (ms * 1e-3) / sms / 1.9e9);
        ms = timeit([&] { k_mufu_ffma<4><<<grid, T>>>(out, 0.5f); });
        printf("[ctas/sm %d] MUFU+4FFMA: %.3f ms -> mufu %.3f /clk/SM, ffma %.2f /clk/SM\n", wps, ms, warps * ITERS * 8 / (ms * 1e-3) / sms / 1.9e9, warps * ITERS * 32 / (ms * 1e-3) / sms / 1.9e9);
        ms = timeit([&] { k_mufu_ffma<8><<<grid, T>>>(out, 0.5f); });
        printf("[ctas/sm %d] MUFU+8FFMA: %.3f ms -> mufu %.3f /clk/SM, ffma %.2f /clk/SM\n", wps, ms, warps * ITERS * 8 / (ms * 1e-3) / sms / 1.9e9, warps * ITERS * 64 / (ms * 1e-3) / sms / 1.9e9);
        ms = timeit([&] { k_dfma<<<grid, T>>>((double*)out, 0.999); });
        printf("[ctas/sm %d] DFMA   : %.3f ms -> %.3f warp-instr/clk/SM\n", wps, ms, warps * ITERS * 8 / (ms * 1e-3) / sms / 1.9e9);
    }
    // report actual SM clock while busy
    printf("done\n");
    return 0;
}
