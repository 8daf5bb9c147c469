// microbench.cu -- issue rates of the instructions the FFT engine is made of (B200, sm_100a).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o microbench microbench.cu && ./microbench
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048
template <int MODE>
__global__ void k_fp(float2* out, float2 seed) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(seed.x + i + threadIdx.x, seed.y - i);
    const float2 b = make_float2(seed.y, seed.x), c = make_float2(0.5f, 0.25f);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) a[i] = __fadd2_rn(a[i], b);
            if (MODE == 1) a[i] = __ffma2_rn(a[i], b, c);
            if (MODE == 2) a[i] = __fmul2_rn(a[i], b);
            if (MODE == 3) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }   // 2 scalar FFMA
            if (MODE == 4) { a[i].x = a[i].x + b.x; a[i].y = a[i].y + b.y; }                      // 2 scalar FADD
            if (MODE == 5) a[i] = __fmul2_rn(a[i], make_float2(b.x, b.x));                           // broadcast scalar operand
        }
    }
    float2 s = a[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) { s.x += a[i].x; s.y += a[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int BYTES>
__global__ void k_smem(float* out) {
    extern __shared__ float4 sm[];
    const int tid = threadIdx.x;
    float acc = 0.f;
    for (int i = tid; i < 4096; i += blockDim.x) sm[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    for (int it = 0; it < 512; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int idx = (tid + r * 256 + it) & 4095;
            if (BYTES == 16) { float4 v = sm[idx]; acc += v.x + v.w; }
            if (BYTES == 8) { float2 v = reinterpret_cast<float2*>(sm)[idx]; acc += v.x + v.y; }
            if (BYTES == 4) { float v = reinterpret_cast<float*>(sm)[idx]; acc += v; }
        }
    }
    out[blockIdx.x * blockDim.x + tid] = acc;
}

template <class F>
float time_ms(F f) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < 5; ++i) f();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / 5;
}

int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int sms = pr.multiProcessorCount; int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    printf("%s, %d SMs, nominal %d MHz\n", pr.name, sms, khz / 1000);
    float2* out; cudaMalloc(&out, sizeof(float2) * sms * 8 * 1024);
    const char* names[] = {"FADD2", "FFMA2", "FMUL2", "2xFFMA scalar", "2xFADD scalar", "FMUL2 bcast"};
    for (int warps = 4; warps <= 32; warps *= 2) {
        const int threads = warps * 32;
        auto run = [&](int mode) {
            float ms = 0;
            switch (mode) {
                case 0: ms = time_ms([&] { k_fp<0><<<sms, threads>>>(out, make_float2(1, 2)); }); break;
                case 1: ms = time_ms([&] { k_fp<1><<<sms, threads>>>(out, make_float2(1, 2)); }); break;
                case 2: ms = time_ms([&] { k_fp<2><<<sms, threads>>>(out, make_float2(1, 2)); }); break;
                case 3: ms = time_ms([&] { k_fp<3><<<sms, threads>>>(out, make_float2(1, 2)); }); break;
                case 4: ms = time_ms([&] { k_fp<4><<<sms, threads>>>(out, make_float2(1, 2)); }); break;
                case 5: ms = time_ms([&] { k_fp<5><<<sms, threads>>>(out, make_float2(1, 2)); }); break;
            }
            // "packed-op equivalents" per SM per ns
            double ops = (double)ITERS * 8 * warps;  // warp-level packed ops (or scalar pairs) per SM
            printf("  warps/SM %2d %-14s %.3f ms  -> %.2f warp-ops/SM/ns (x1.9 GHz => %.2f per clk/SM)\n", warps, names[mode], ms, ops / (ms * 1e6), ops / (ms * 1e6) / 1.9);
        };
        for (int m = 0; m < 6; ++m) run(m);
    }
    cudaFuncSetAttribute(k_smem<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k_smem<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    cudaFuncSetAttribute(k_smem<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int threads = 128; threads <= 512; threads *= 2) {
        float m16 = time_ms([&] { k_smem<16><<<sms, threads, 65536>>>((float*)out); });
        float m8 = time_ms([&] { k_smem<8><<<sms, threads, 65536>>>((float*)out); });
        float m4 = time_ms([&] { k_smem<4><<<sms, threads, 65536>>>((float*)out); });
        double n = 512.0 * 8 * threads;  // loads per SM
        printf("  smem threads %3d: LDS.128 %.1f B/ns/SM  LDS.64 %.1f B/ns/SM  LDS.32 %.1f B/ns/SM\n", threads, n * 16 / (m16 * 1e6), n * 8 / (m8 * 1e6), n * 4 / (m4 * 1e6));
    }
    return 0;
}
