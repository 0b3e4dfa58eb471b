// Pipe-rate micro-benchmark for sm_100a: warp-instructions per clock per SM for the instruction classes the colour stage
// leans on (scalar and packed FP32, IMAD, ALU ops, and mixes).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench tools/ubench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define REP 64
#define ITERS 256

template <int MODE>
__global__ void __launch_bounds__(1024, 1) kern(float *out, unsigned long long *cyc, float a, float b, unsigned ia, unsigned ib) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    unsigned i0 = threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3, j0 = i0 * 3, j1 = i0 * 5, j2 = i0 * 7, j3 = i0 * 9;
    unsigned long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int r = 0; r < REP / 8; ++r) {
            if (MODE == 0) {  // 8 scalar FFMA
                asm volatile("fma.rn.f32 %0,%0,%8,%9; fma.rn.f32 %1,%1,%8,%9; fma.rn.f32 %2,%2,%8,%9; fma.rn.f32 %3,%3,%8,%9;"
                             "fma.rn.f32 %4,%4,%8,%9; fma.rn.f32 %5,%5,%8,%9; fma.rn.f32 %6,%6,%8,%9; fma.rn.f32 %7,%7,%8,%9;"
                             : "+f"(x0), "+f"(x1), "+f"(x2), "+f"(x3), "+f"(x4), "+f"(x5), "+f"(x6), "+f"(x7) : "f"(a), "f"(b));
            } else if (MODE == 1) {  // 8 packed FFMA2 (16 flops-pairs)
                asm volatile("{.reg .b64 p,q,r,s,m,n; mov.b64 m,{%8,%8}; mov.b64 n,{%9,%9};"
                             "mov.b64 p,{%0,%1}; mov.b64 q,{%2,%3}; mov.b64 r,{%4,%5}; mov.b64 s,{%6,%7};"
                             "fma.rn.f32x2 p,p,m,n; fma.rn.f32x2 q,q,m,n; fma.rn.f32x2 r,r,m,n; fma.rn.f32x2 s,s,m,n;"
                             "fma.rn.f32x2 p,p,m,n; fma.rn.f32x2 q,q,m,n; fma.rn.f32x2 r,r,m,n; fma.rn.f32x2 s,s,m,n;"
                             "mov.b64 {%0,%1},p; mov.b64 {%2,%3},q; mov.b64 {%4,%5},r; mov.b64 {%6,%7},s;}"
                             : "+f"(x0), "+f"(x1), "+f"(x2), "+f"(x3), "+f"(x4), "+f"(x5), "+f"(x6), "+f"(x7) : "f"(a), "f"(b));
            } else if (MODE == 2) {  // 8 IMAD
                asm volatile("mad.lo.u32 %0,%0,%4,%5; mad.lo.u32 %1,%1,%4,%5; mad.lo.u32 %2,%2,%4,%5; mad.lo.u32 %3,%3,%4,%5;"
                             "mad.lo.u32 %0,%0,%4,%5; mad.lo.u32 %1,%1,%4,%5; mad.lo.u32 %2,%2,%4,%5; mad.lo.u32 %3,%3,%4,%5;"
                             : "+r"(i0), "+r"(i1), "+r"(i2), "+r"(i3) : "r"(ia), "r"(ib));
            } else if (MODE == 3) {  // 8 LOP3 (xor)
                asm volatile("lop3.b32 %0,%0,%4,%5,0x96; lop3.b32 %1,%1,%4,%5,0x96; lop3.b32 %2,%2,%4,%5,0x96; lop3.b32 %3,%3,%4,%5,0x96;"
                             "lop3.b32 %0,%0,%5,%4,0x96; lop3.b32 %1,%1,%5,%4,0x96; lop3.b32 %2,%2,%5,%4,0x96; lop3.b32 %3,%3,%5,%4,0x96;"
                             : "+r"(i0), "+r"(i1), "+r"(i2), "+r"(i3) : "r"(ia), "r"(ib));
            } else if (MODE == 4) {  // 4 FFMA2 + 4 LOP3
                asm volatile("{.reg .b64 p,q,r,s,m,n; mov.b64 m,{%12,%12}; mov.b64 n,{%13,%13};"
                             "mov.b64 p,{%0,%1}; mov.b64 q,{%2,%3}; mov.b64 r,{%4,%5}; mov.b64 s,{%6,%7};"
                             "fma.rn.f32x2 p,p,m,n; lop3.b32 %8,%8,%14,%15,0x96; fma.rn.f32x2 q,q,m,n; lop3.b32 %9,%9,%14,%15,0x96;"
                             "fma.rn.f32x2 r,r,m,n; lop3.b32 %10,%10,%14,%15,0x96; fma.rn.f32x2 s,s,m,n; lop3.b32 %11,%11,%14,%15,0x96;"
                             "mov.b64 {%0,%1},p; mov.b64 {%2,%3},q; mov.b64 {%4,%5},r; mov.b64 {%6,%7},s;}"
                             : "+f"(x0), "+f"(x1), "+f"(x2), "+f"(x3), "+f"(x4), "+f"(x5), "+f"(x6), "+f"(x7), "+r"(i0), "+r"(i1), "+r"(i2), "+r"(i3)
                             : "f"(a), "f"(b), "r"(ia), "r"(ib));
            } else if (MODE == 5) {  // 4 FFMA + 4 LOP3
                asm volatile("fma.rn.f32 %0,%0,%8,%9; lop3.b32 %4,%4,%10,%11,0x96; fma.rn.f32 %1,%1,%8,%9; lop3.b32 %5,%5,%10,%11,0x96;"
                             "fma.rn.f32 %2,%2,%8,%9; lop3.b32 %6,%6,%10,%11,0x96; fma.rn.f32 %3,%3,%8,%9; lop3.b32 %7,%7,%10,%11,0x96;"
                             : "+f"(x0), "+f"(x1), "+f"(x2), "+f"(x3), "+r"(i0), "+r"(i1), "+r"(i2), "+r"(i3) : "f"(a), "f"(b), "r"(ia), "r"(ib));
            } else if (MODE == 6) {  // 4 IMAD + 4 LOP3
                asm volatile("mad.lo.u32 %0,%0,%8,%9; lop3.b32 %4,%4,%8,%9,0x96; mad.lo.u32 %1,%1,%8,%9; lop3.b32 %5,%5,%8,%9,0x96;"
                             "mad.lo.u32 %2,%2,%8,%9; lop3.b32 %6,%6,%8,%9,0x96; mad.lo.u32 %3,%3,%8,%9; lop3.b32 %7,%7,%8,%9,0x96;"
                             : "+r"(i0), "+r"(i1), "+r"(i2), "+r"(i3), "+r"(j0), "+r"(j1), "+r"(j2), "+r"(j3) : "r"(ia), "r"(ib));
            } else if (MODE == 7) {  // 4 IMAD + 4 FFMA
                asm volatile("mad.lo.u32 %4,%4,%10,%11; fma.rn.f32 %0,%0,%8,%9; mad.lo.u32 %5,%5,%10,%11; fma.rn.f32 %1,%1,%8,%9;"
                             "mad.lo.u32 %6,%6,%10,%11; fma.rn.f32 %2,%2,%8,%9; mad.lo.u32 %7,%7,%10,%11; fma.rn.f32 %3,%3,%8,%9;"
                             : "+f"(x0), "+f"(x1), "+f"(x2), "+f"(x3), "+r"(i0), "+r"(i1), "+r"(i2), "+r"(i3) : "f"(a), "f"(b), "r"(ia), "r"(ib));
            } else if (MODE == 10) {  // 4 IMAD + 4 FFMA2
                asm volatile("{.reg .b64 p,q,r,s,m,n; mov.b64 m,{%12,%12}; mov.b64 n,{%13,%13};"
                             "mov.b64 p,{%0,%1}; mov.b64 q,{%2,%3}; mov.b64 r,{%4,%5}; mov.b64 s,{%6,%7};"
                             "fma.rn.f32x2 p,p,m,n; mad.lo.u32 %8,%8,%14,%15; fma.rn.f32x2 q,q,m,n; mad.lo.u32 %9,%9,%14,%15;"
                             "fma.rn.f32x2 r,r,m,n; mad.lo.u32 %10,%10,%14,%15; fma.rn.f32x2 s,s,m,n; mad.lo.u32 %11,%11,%14,%15;"
                             "mov.b64 {%0,%1},p; mov.b64 {%2,%3},q; mov.b64 {%4,%5},r; mov.b64 {%6,%7},s;}"
                             : "+f"(x0), "+f"(x1), "+f"(x2), "+f"(x3), "+f"(x4), "+f"(x5), "+f"(x6), "+f"(x7), "+r"(i0), "+r"(i1), "+r"(i2), "+r"(i3)
                             : "f"(a), "f"(b), "r"(ia), "r"(ib));
            } else if (MODE == 11) {  // 4 IMAD + 8 FFMA (the same float work as mode 10, scalar) -> counted as 12 instructions per 8
                asm volatile("fma.rn.f32 %0,%0,%12,%13; fma.rn.f32 %1,%1,%12,%13; mad.lo.u32 %8,%8,%14,%15; fma.rn.f32 %2,%2,%12,%13; fma.rn.f32 %3,%3,%12,%13; mad.lo.u32 %9,%9,%14,%15;"
                             "fma.rn.f32 %4,%4,%12,%13; fma.rn.f32 %5,%5,%12,%13; mad.lo.u32 %10,%10,%14,%15; fma.rn.f32 %6,%6,%12,%13; fma.rn.f32 %7,%7,%12,%13; mad.lo.u32 %11,%11,%14,%15;"
                             : "+f"(x0), "+f"(x1), "+f"(x2), "+f"(x3), "+f"(x4), "+f"(x5), "+f"(x6), "+f"(x7), "+r"(i0), "+r"(i1), "+r"(i2), "+r"(i3)
                             : "f"(a), "f"(b), "r"(ia), "r"(ib));
            } else if (MODE == 12) {  // 2 IMAD + 2 LOP3 + 2 FFMA2 + 2 more LOP3 (kernel-like mix)
                asm volatile("{.reg .b64 p,q,m,n; mov.b64 m,{%12,%12}; mov.b64 n,{%13,%13}; mov.b64 p,{%0,%1}; mov.b64 q,{%2,%3};"
                             "fma.rn.f32x2 p,p,m,n; mad.lo.u32 %8,%8,%14,%15; lop3.b32 %10,%10,%14,%15,0x96; lop3.b32 %11,%11,%14,%15,0x96;"
                             "fma.rn.f32x2 q,q,m,n; mad.lo.u32 %9,%9,%14,%15; lop3.b32 %10,%10,%15,%14,0x96; lop3.b32 %11,%11,%15,%14,0x96;"
                             "mov.b64 {%0,%1},p; mov.b64 {%2,%3},q;}"
                             : "+f"(x0), "+f"(x1), "+f"(x2), "+f"(x3), "+f"(x4), "+f"(x5), "+f"(x6), "+f"(x7), "+r"(i0), "+r"(i1), "+r"(i2), "+r"(i3)
                             : "f"(a), "f"(b), "r"(ia), "r"(ib));
            } else if (MODE == 13) {  // 8 FFMA with distinct constant operands per instruction pair (fewer register-bank clashes)
                asm volatile("fma.rn.f32 %0,%0,%8,%9; fma.rn.f32 %1,%1,%9,%8; fma.rn.f32 %2,%2,%8,%9; fma.rn.f32 %3,%3,%9,%8;"
                             "fma.rn.f32 %4,%4,%8,%9; fma.rn.f32 %5,%5,%9,%8; fma.rn.f32 %6,%6,%8,%9; fma.rn.f32 %7,%7,%9,%8;"
                             : "+f"(x0), "+f"(x1), "+f"(x2), "+f"(x3), "+f"(x4), "+f"(x5), "+f"(x6), "+f"(x7) : "f"(a), "f"(b));
            } else if (MODE == 14) {  // 8 IMAD with an immediate multiplier
                asm volatile("mad.lo.u32 %0,%0,12345,%4; mad.lo.u32 %1,%1,12345,%4; mad.lo.u32 %2,%2,12345,%4; mad.lo.u32 %3,%3,12345,%4;"
                             "mad.lo.u32 %0,%0,54321,%4; mad.lo.u32 %1,%1,54321,%4; mad.lo.u32 %2,%2,54321,%4; mad.lo.u32 %3,%3,54321,%4;"
                             : "+r"(i0), "+r"(i1), "+r"(i2), "+r"(i3) : "r"(ia));
            } else if (MODE == 15) {  // 8 IADD3
                asm volatile("add.u32 %0,%0,%4; add.u32 %1,%1,%4; add.u32 %2,%2,%4; add.u32 %3,%3,%4;"
                             "add.u32 %0,%0,%5; add.u32 %1,%1,%5; add.u32 %2,%2,%5; add.u32 %3,%3,%5;"
                             : "+r"(i0), "+r"(i1), "+r"(i2), "+r"(i3) : "r"(ia), "r"(ib));
            } else if (MODE == 8) {  // 8 FADD (scalar)
                asm volatile("add.rn.f32 %0,%0,%8; add.rn.f32 %1,%1,%8; add.rn.f32 %2,%2,%8; add.rn.f32 %3,%3,%8;"
                             "add.rn.f32 %4,%4,%8; add.rn.f32 %5,%5,%8; add.rn.f32 %6,%6,%8; add.rn.f32 %7,%7,%8;"
                             : "+f"(x0), "+f"(x1), "+f"(x2), "+f"(x3), "+f"(x4), "+f"(x5), "+f"(x6), "+f"(x7) : "f"(a));
            } else if (MODE == 9) {  // 8 VIMNMX (min)
                asm volatile("min.s32 %0,%0,%4; min.s32 %1,%1,%4; min.s32 %2,%2,%4; min.s32 %3,%3,%4;"
                             "max.s32 %0,%0,%5; max.s32 %1,%1,%5; max.s32 %2,%2,%5; max.s32 %3,%3,%5;"
                             : "+r"(i0), "+r"(i1), "+r"(i2), "+r"(i3) : "r"(ia), "r"(ib));
            }
        }
    }
    unsigned long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + (float)(i0 + i1 + i2 + i3 + j0 + j1 + j2 + j3);
}

template <int MODE>
void run(const char *name) {
    const int blocks = 148;
    float *out;
    unsigned long long *cyc, h[148];
    cudaMalloc(&out, blocks * 1024 * sizeof(float));
    cudaMalloc(&cyc, blocks * sizeof(unsigned long long));
    kern<MODE><<<blocks, 1024>>>(out, cyc, 1.0001f, 0.5f, 3u, 5u);
    kern<MODE><<<blocks, 1024>>>(out, cyc, 1.0001f, 0.5f, 3u, 5u);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < blocks; ++i) avg += (double)h[i];
    avg /= blocks;
    const double winstr = 32.0 * ITERS * REP;  // warp-instructions per SM (32 warps)
    printf("%-28s %8.0f cycles  %.3f warp-instr/clk/SM\n", name, avg, winstr / avg);
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    run<0>("FFMA");
    run<1>("FFMA2");
    run<2>("IMAD");
    run<3>("LOP3");
    run<4>("FFMA2+LOP3 1:1");
    run<5>("FFMA+LOP3 1:1");
    run<6>("IMAD+LOP3 1:1");
    run<7>("IMAD+FFMA 1:1");
    run<8>("FADD");
    run<9>("VIMNMX");
    run<10>("IMAD+FFMA2 1:1");
    run<11>("IMAD+FFMA 1:2 (x1.5 instr)");
    run<12>("FFMA2+IMAD+2xLOP3");
    run<13>("FFMA alt operands");
    run<14>("IMAD imm");
    run<15>("IADD");
    return 0;
}
