// ubench.cu -- sm_100a pipe-throughput microbenchmarks (design aid; measures the integer / FP64 roofline
// denominators used in DESIGN.md).  Each kernel issues ITER x 8 independent ops per thread.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITER = 2048;

#define KERNEL(name, DECL, BODY, SINK)                                                    \
__global__ void __launch_bounds__(256) name(unsigned long long *out, unsigned seed) {     \
    DECL;                                                                                 \
    long long t0 = clock64();                                                             \
    for (int it = 0; it < ITER; it++) { BODY; }                                           \
    long long t1 = clock64();                                                             \
    unsigned long long s = SINK;                                                          \
    if ((unsigned)s == 0x12345678u) out[1] = s;                                                  \
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);      \
}

#define R8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)
#define DECL32 unsigned a0=seed,a1=seed+1,a2=seed+2,a3=seed+3,a4=seed+4,a5=seed+5,a6=seed+6,a7=seed+7, b=threadIdx.x|1, c=seed*3+1
#define SINK32 (unsigned long long)(a0^a1^a2^a3^a4^a5^a6^a7)
#define IMADLO(i) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a##i) : "r"(b), "r"(c));
#define IMADHI(i) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a##i) : "r"(b), "r"(c));
#define IADD(i) asm volatile("add.u32 %0, %0, %1;" : "+r"(a##i) : "r"(b));
#define LOP(i) asm volatile("xor.b32 %0, %0, %1;" : "+r"(a##i) : "r"(b));
#define IADD3x(i) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(a##i) : "r"(b), "r"(c));
KERNEL(k_imad_lo, DECL32, R8(IMADLO), SINK32)
KERNEL(k_imad_hi, DECL32, R8(IMADHI), SINK32)
KERNEL(k_iadd, DECL32, R8(IADD), SINK32)
KERNEL(k_lop, DECL32, R8(LOP), SINK32)

#define DECL64 unsigned long long a0=seed,a1=seed+1,a2=seed+2,a3=seed+3,a4=seed+4,a5=seed+5,a6=seed+6,a7=seed+7; unsigned b=threadIdx.x|1, c=seed*3+1
#define SINK64 (a0^a1^a2^a3^a4^a5^a6^a7)
#define IMADWIDE(i) asm volatile("{.reg .u32 lo; cvt.u32.u64 lo, %0; mad.wide.u32 %0, lo, %1, %0;}" : "+l"(a##i) : "r"(b));
#define MULHI64(i) asm volatile("mul.hi.u64 %0, %0, %1;" : "+l"(a##i) : "l"(bb));
#define MULLO64(i) asm volatile("mul.lo.u64 %0, %0, %1;" : "+l"(a##i) : "l"(bb));
#define ADD64(i) asm volatile("add.u64 %0, %0, %1;" : "+l"(a##i) : "l"(bb));
KERNEL(k_imad_wide, DECL64, R8(IMADWIDE), SINK64)
KERNEL(k_mulhi64, DECL64; unsigned long long bb = 0x9e3779b97f4a7c15ull + threadIdx.x, R8(MULHI64), SINK64)
KERNEL(k_mullo64, DECL64; unsigned long long bb = 0x9e3779b97f4a7c15ull + threadIdx.x, R8(MULLO64), SINK64)
KERNEL(k_add64, DECL64; unsigned long long bb = 0x9e3779b97f4a7c15ull + threadIdx.x, R8(ADD64), SINK64)

#define DECLD double a0=seed,a1=seed+1,a2=seed+2,a3=seed+3,a4=seed+4,a5=seed+5,a6=seed+6,a7=seed+7, b=1.0000001+threadIdx.x*1e-9, c=1e-7
#define SINKD (unsigned long long)__double_as_longlong(a0+a1+a2+a3+a4+a5+a6+a7)
#define DFMA(i) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a##i) : "d"(b), "d"(c));
#define DMUL(i) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(a##i) : "d"(b));
#define DADD(i) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a##i) : "d"(c));
KERNEL(k_dfma, DECLD, R8(DFMA), SINKD)
KERNEL(k_dmul, DECLD, R8(DMUL), SINKD)
KERNEL(k_dadd, DECLD, R8(DADD), SINKD)

#define DECLF float a0=seed,a1=seed+1,a2=seed+2,a3=seed+3,a4=seed+4,a5=seed+5,a6=seed+6,a7=seed+7, b=1.0000001f, c=1e-7f
#define SINKF (unsigned long long)__float_as_uint(a0+a1+a2+a3+a4+a5+a6+a7)
#define FFMA(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a##i) : "f"(b), "f"(c));
KERNEL(k_ffma, DECLF, R8(FFMA), SINKF)

// mixed: 4 IMAD.lo + 4 DFMA per iteration-slot (do the two pipes overlap?)
#define DECLMIX unsigned a0=seed,a1=seed+1,a2=seed+2,a3=seed+3, b=threadIdx.x|1, c=seed*3+1; double d0=seed,d1=seed+1,d2=seed+2,d3=seed+3, e=1.0000001, f=1e-7
#define MIXBODY IMADLO(0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d0) : "d"(e), "d"(f)); IMADLO(1) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d1) : "d"(e), "d"(f)); IMADLO(2) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d2) : "d"(e), "d"(f)); IMADLO(3) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d3) : "d"(e), "d"(f));
KERNEL(k_mix_imad_dfma, DECLMIX, MIXBODY, (unsigned long long)(a0^a1^a2^a3) + (unsigned long long)__double_as_longlong(d0+d1+d2+d3))
// mixed: 4 IMAD.lo + 4 IADD (fma pipe + alu pipe)
#define MIX2 IMADLO(0) IADD(4) IMADLO(1) IADD(5) IMADLO(2) IADD(6) IMADLO(3) IADD(7)
KERNEL(k_mix_imad_iadd, DECL32, MIX2, SINK32)
// mixed: 4 IMAD.WIDE + 4 IADD
#define DECLMIX3 unsigned long long a0=seed,a1=seed+1,a2=seed+2,a3=seed+3; unsigned a4=seed+4,a5=seed+5,a6=seed+6,a7=seed+7, b=threadIdx.x|1, c=seed*3+1
#define MIX3 IMADWIDE(0) IADD(4) IMADWIDE(1) IADD(5) IMADWIDE(2) IADD(6) IMADWIDE(3) IADD(7)
KERNEL(k_mix_wide_iadd, DECLMIX3, MIX3, (a0^a1^a2^a3) + (a4^a5^a6^a7))
// conversions
#define DECLCV unsigned long long a0=seed,a1=seed+1,a2=seed+2,a3=seed+3,a4=seed+4,a5=seed+5,a6=seed+6,a7=seed+7
#define CVT(i) asm volatile("{.reg .f64 t; cvt.rn.f64.u64 t, %0; cvt.rzi.u64.f64 %0, t;}" : "+l"(a##i));
KERNEL(k_cvt_u64_f64_roundtrip, DECLCV, R8(CVT), SINK64)

template <typename K> int run(const char *name, K k, int ops_per_iter_thread, unsigned long long *d) {
    int blocks = 148 * 8, threads = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<blocks, threads>>>(d, 12345u);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    k<<<blocks, threads>>>(d, 12345u);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long cyc; cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    double ops = (double)blocks * threads * ITER * ops_per_iter_thread;
    // all 148*8 CTAs are co-resident (one wave), so block 0's cycle count ~ kernel duration in SM clocks
    printf("%-28s %8.3f ms  %10.3e ops/s  %7.1f ops/clk/SM  sm_clock=%.0f MHz\n", name, ms, ops / (ms * 1e-3),
           ops / ((double)cyc * 148.0), (double)cyc / (ms * 1e3));
    return 0;
}

int main() {
    unsigned long long *d; CK(cudaMalloc(&d, 64));
    run("imad.lo.u32", k_imad_lo, 8, d);
    run("imad.hi.u32", k_imad_hi, 8, d);
    run("imad.wide.u32", k_imad_wide, 8, d);
    run("iadd.u32", k_iadd, 8, d);
    run("lop(xor).b32", k_lop, 8, d);
    run("mul.hi.u64", k_mulhi64, 8, d);
    run("mul.lo.u64", k_mullo64, 8, d);
    run("add.u64", k_add64, 8, d);
    run("dfma", k_dfma, 8, d);
    run("dmul", k_dmul, 8, d);
    run("dadd", k_dadd, 8, d);
    run("ffma", k_ffma, 8, d);
    run("mix 4imad+4dfma", k_mix_imad_dfma, 8, d);
    run("mix 4imad+4iadd", k_mix_imad_iadd, 8, d);
    run("mix 4wide+4iadd", k_mix_wide_iadd, 8, d);
    run("cvt u64->f64->u64 (2 cvt)", k_cvt_u64_f64_roundtrip, 16, d);
    return 0;
}
