// Throughput of the MUFU flavours an epilogue can use for GELU / sigmoid (per SM and clock), B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_probe mufu_probe.cu && ./mufu_probe
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int MODE>
__global__ void k(float* out, int iters) {
    float a = threadIdx.x * 1e-3f, b = a + 0.1f, c = a + 0.2f, d = a + 0.3f;
    uint32_t ha = threadIdx.x * 17u + 0x3000u, hb = ha + 5, hc = ha + 9, hd = ha + 11;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {
            asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(b));
            asm volatile("tanh.approx.f32 %0, %0;" : "+f"(c)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(d));
        } else if (MODE == 1) {
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(c)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(d));
        } else if (MODE == 2) {
            asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(ha)); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(hb));
            asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(hc)); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(hd));
        } else if (MODE == 3) {
            asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(ha)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(hb));
            asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(hc)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(hd));
        } else if (MODE == 4) {
            asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(b));
            asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(c)); asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(d));
        } else if (MODE == 5) {
            asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(ha)); asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(hb));
            asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(hc)); asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(hd));
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d + __uint_as_float(ha ^ hb ^ hc ^ hd);
    if (threadIdx.x == 0 && blockIdx.x == 0) reinterpret_cast<long long*>(out)[4096] = t1 - t0;
}

template <int MODE> void run(const char* name, int elems_per_instr) {
    float* out;
    cudaMalloc(&out, 1 << 22);
    const int iters = 4096, threads = 512;          // 16 warps per SM = 4 per scheduler
    k<MODE><<<148, threads>>>(out, iters);
    k<MODE><<<148, threads>>>(out, iters);
    cudaDeviceSynchronize();
    long long cyc;
    cudaMemcpy(&cyc, reinterpret_cast<long long*>(out) + 4096, 8, cudaMemcpyDeviceToHost);
    const double instr = (double)iters * 4 * threads;          // thread-instructions per SM
    printf("%-22s %8.2f results / clk / SM   (%.2f thread-instr / clk / SM)\n", name, instr * elems_per_instr / cyc, instr / cyc);
    cudaFree(out);
}

int main() {
    run<0>("tanh.approx.f32", 1);
    run<1>("ex2.approx.ftz.f32", 1);
    run<2>("tanh.approx.f16x2", 2);
    run<3>("ex2.approx.f16x2", 2);
    run<4>("rcp.approx.ftz.f32", 1);
    run<5>("tanh.approx.bf16x2", 2);
    return 0;
}
