// Microbenchmark: issue throughput of packed fp32 (FFMA2/FMUL2/FADD2, PTX *.f32x2) vs scalar FFMA on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o f32x2 f32x2.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float2 upk(u64 r) { float2 d; asm("mov.b64 {%0,%1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(r)); return d; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
constexpr int ILP = 8, ITERS = 4096;
template <int MODE>
__global__ void k(float* out, float s) {
    float a[ILP], b[ILP];
    u64 p[ILP];
    for (int i = 0; i < ILP; i++) { a[i] = threadIdx.x * 1e-3f + i; b[i] = a[i] * 0.5f; p[i] = pk(a[i], b[i]); }
    const u64 ps = pk(s, s), pc = pk(1e-6f, 1e-6f);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (MODE == 0) { a[i] = __fmaf_rn(a[i], s, 1e-6f); b[i] = __fmaf_rn(b[i], s, 1e-6f); }   // 2 scalar FFMA
            if (MODE == 1) p[i] = fma2(p[i], ps, pc);                                               // 1 FFMA2
            if (MODE == 2) { a[i] = __fmul_rn(a[i], s); b[i] = __fadd_rn(b[i], s); }                // FMUL + FADD
            if (MODE == 3) { p[i] = mul2(p[i], ps); p[i] = add2(p[i], pc); }                        // FMUL2 + FADD2 (4 flop-pairs)
            if (MODE == 4) { p[i] = fma2(p[i], ps, pc); a[i] = __fmaf_rn(a[i], s, 1e-6f); }         // mixed
            if (MODE == 5) p[i] = fma2(p[i], pk(0.999f, 0.999f), pk(1e-6f, 1e-6f));                  // FFMA2 with immediates
            if (MODE == 6) p[i] = fma2(p[i], ps, p[i]);                                             // FFMA2, 2 distinct register pairs
            if (MODE == 7) { p[i] = mul2(p[i], ps); a[i] = __fmul_rn(a[i], s); }                    // FMUL2 + FMUL
        }
    }
    float acc = 0;
    for (int i = 0; i < ILP; i++) { float2 q = upk(p[i]); acc += a[i] + b[i] + q.x + q.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE>
void run(const char* name, double lane_ops_per_iter) {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(out, 0.999f);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) k<MODE><<<148 * 8, 256>>>(out, 0.999f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    const double thr = 148.0 * 8 * 256;
    printf("%-28s %.3f ms  %.1f G lane-fp-ops/s  (%.1f G warp-instr/s)\n", name, ms, thr * ITERS * ILP * lane_ops_per_iter / ms / 1e6,
           thr / 32 * ITERS * ILP * (MODE == 0 || MODE == 2 || MODE == 3 || MODE == 4 || MODE == 7 ? 2 : 1) / ms / 1e6);
    cudaFree(out);
}
int main() {
    run<0>("2x FFMA (scalar)", 2);
    run<1>("1x FFMA2 (packed)", 2);
    run<2>("FMUL + FADD (scalar)", 2);
    run<3>("FMUL2 + FADD2 (packed)", 4);
    run<4>("FFMA2 + FFMA (mixed)", 3);
    run<5>("1x FFMA2 (immediate b, c)", 2);
    run<6>("1x FFMA2 (uniform b, c = a)", 2);
    run<7>("FMUL2 + FMUL (mixed)", 3);
    return 0;
}
