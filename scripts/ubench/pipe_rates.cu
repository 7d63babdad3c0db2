// Issue rate of the instructions a BVH8 node test can be built from, on one SM sub-partition (B200, sm_100a).
// Every kernel runs 8 independent dependency chains per thread, 16 warps per SM (4 per scheduler), and reports
// warp instructions per cycle per scheduler from clock64().   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
#define ITER 4096
#define CH 8
template <int OP>
__global__ void __launch_bounds__(512) k(uint32_t* out, long long* cyc, uint32_t seed) {
    uint32_t r[CH];
    float f[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { r[i] = seed * (threadIdx.x + 1) + i * 0x01010101u; f[i] = __uint_as_float(0x3f800000u | (r[i] & 0xffffu)); }
    const float fa = __uint_as_float(0x3f800001u + seed), fb = __uint_as_float(0x3a800000u + seed);
    const uint32_t ua = 0x3c003c01u + seed, ub = 0x64646464u ^ seed;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (OP == 0) f[i] = fmaf(f[i], fa, fb);                                         // FFMA
            if (OP == 1) { __half2 h = *reinterpret_cast<__half2*>(&r[i]); h = __hfma2(h, *reinterpret_cast<const __half2*>(&ua), *reinterpret_cast<const __half2*>(&ub)); r[i] = *reinterpret_cast<uint32_t*>(&h); }   // HFMA2
            if (OP == 2) { float d; asm volatile("{.reg .f16 l_, h_; mov.b32 {l_, h_}, %1; fma.rn.f32.f16 %0, l_, h_, %2;}" : "=f"(d) : "r"(r[i]), "f"(f[i])); f[i] = d; }   // FHFMA
            if (OP == 3) r[i] = __byte_perm(r[i], ub, 0x4140 + (seed & 1));                 // PRMT
            if (OP == 4) f[i] = fmaxf(f[i], fa);                                            // FMNMX
            if (OP == 5) f[i] = fmaxf(fmaxf(f[i], fa), fb);                                 // FMNMX3
            if (OP == 6) { __half2 h = *reinterpret_cast<__half2*>(&r[i]); h = __hmax2(h, *reinterpret_cast<const __half2*>(&ua)); r[i] = *reinterpret_cast<uint32_t*>(&h); }   // HMNMX2
            if (OP == 7) { __half2 h = *reinterpret_cast<__half2*>(&r[i]); r[i] += __hgt2_mask(h, *reinterpret_cast<const __half2*>(&ua)); }   // HSET2 / HSETP2 + ...
            if (OP == 8) { __half2 h = *reinterpret_cast<__half2*>(&r[i]); f[i] = __low2float(h) + f[i]; r[i] += 0x00010001u; }   // HADD2.F32 (+ FADD + IADD)
            if (OP == 9) r[i] = (r[i] & ua) ^ ub;                                           // LOP3
            if (OP == 10) r[i] = (f[i] > fa) ? r[i] : ua;                                   // FSETP + SEL (f constant: hoisted?) 
            if (OP == 11) r[i] = __umulhi(r[i], 256u) + 0x4b000000u;                        // IMAD.HI
            if (OP == 12) { __half2 h = *reinterpret_cast<__half2*>(&r[i]); h = __hadd2(h, *reinterpret_cast<const __half2*>(&ua)); r[i] = *reinterpret_cast<uint32_t*>(&h); }   // HADD2
            if (OP == 13) { __half2 h = *reinterpret_cast<__half2*>(&r[i]); h = __hmax2(__hmax2(h, *reinterpret_cast<const __half2*>(&ua)), *reinterpret_cast<const __half2*>(&ub)); r[i] = *reinterpret_cast<uint32_t*>(&h); }   // 3-input half2 max?
            if (OP == 14) f[i] = (float)(r[i] & 0xffu) + f[i];                              // I2F path
            if (OP == 15) { uint16_t h; asm volatile("cvt.rm.f16.f32 %0, %1;" : "=h"(h) : "f"(f[i])); f[i] += __uint_as_float((uint32_t)h << 13); }   // F2F.F16.F32.RM + SHF + FADD
            if (OP == 16) { uint16_t h; asm volatile("cvt.rn.f16.f32 %0, %1;" : "=h"(h) : "f"(f[i])); f[i] += __uint_as_float((uint32_t)h << 13); }   // F2F.F16.F32 + SHF + FADD
            if (OP == 17) { uint32_t h; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(f[i]), "f"(fa)); f[i] += __uint_as_float(h << 13); }   // F2FP + SHF + FADD
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) acc ^= r[i] ^ __float_as_uint(f[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP> void run(const char* name, uint32_t* out, long long* cyc) {
    k<OP><<<148, 512>>>(out, cyc, 0); cudaDeviceSynchronize();
    k<OP><<<148, 512>>>(out, cyc, 0); cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; i++) c += (double)h[i]; c /= 148;
    // 4 warps per scheduler, each issues ITER*CH "ops"
    printf("%-34s %8.0f cycles  %.3f ops per cycle per scheduler (an op may be more than one SASS instruction: see cuobjdump)\n", name, c, 4.0 * ITER * CH / c);
}
int main() {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
    run<0>("FFMA", out, cyc); run<1>("HFMA2", out, cyc); run<2>("FHFMA (fma.rn.f32.f16)", out, cyc); run<3>("PRMT", out, cyc);
    run<4>("FMNMX", out, cyc); run<5>("FMNMX3 (nested fmaxf)", out, cyc); run<6>("HMNMX2", out, cyc); run<7>("__hgt2_mask + IADD", out, cyc);
    run<8>("HADD2.F32 + FADD + IADD", out, cyc); run<9>("LOP3", out, cyc); run<10>("FSETP + SEL", out, cyc); run<11>("IMAD.HI + IADD", out, cyc);
    run<12>("HADD2", out, cyc); run<13>("nested __hmax2", out, cyc); run<14>("LOP3 + I2F + FADD", out, cyc);
    run<15>("cvt.rm.f16.f32 + SHF + FADD", out, cyc); run<16>("cvt.rn.f16.f32 + SHF + FADD", out, cyc); run<17>("cvt.rn.f16x2.f32 + SHF + FADD", out, cyc);
    return 0;
}
