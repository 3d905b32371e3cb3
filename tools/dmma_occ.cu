// Micro-benchmark: FP64 DMMA (mma.sync m8n8k4) throughput versus resident warps per SM and
// accumulator ILP.  Answers: how many warps per scheduler does the DMMA pipe need?
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_occ dmma_occ.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void k(double* out, int iters) {
    double c[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c[i][0] = i; c[i][1] = i + 0.5; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * (threadIdx.x & 7);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

template <int NACC>
void run(int warps_per_sm, int sms, double* d) {
    int threads = warps_per_sm * 32;
    int blocks = sms;   // one CTA per SM
    if (threads > 1024) { blocks = sms * (threads / 1024); threads = 1024; }
    int iters = 4096 * 32 / NACC;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        cudaEventRecord(e0);
        k<NACC><<<blocks, threads>>>(d, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r) best = ms < best ? ms : best;
    }
    double flops = 512.0 * NACC * iters * (double)blocks * (threads / 32);
    printf("warps/SM %2d  acc/warp %2d : %7.2f TFLOP/s\n", warps_per_sm, NACC, flops / (best * 1e-3) / 1e12);
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* d; cudaMalloc(&d, 64);
    for (int w : {4, 8, 12, 16, 32, 64}) {
        run<4>(w, sms, d); run<8>(w, sms, d); run<16>(w, sms, d); run<32>(w, sms, d);
    }
    return 0;
}
