// Instruction-throughput microbenchmarks for the B200 pipes K3 (split search) leans on.
// Not product code: numbers go into profiles/ and guide the screening arithmetic.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench scripts/microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define ITERS 4096
#define CHAINS 8

template <int OP>
__global__ void __launch_bounds__(512) bench(double *out, const double *in, long long *cycles)
{
    double d[CHAINS];
    float f[CHAINS];
    int n[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) {
        d[k] = in[(threadIdx.x + k) & 63];
        f[k] = (float)d[k];
        n[k] = (int)(d[k] * 1000.0) + k;
    }
    const double a = in[1], b = in[2];
    const float fa = (float)a, fb = (float)b;
    __shared__ double2 sm[1024];
    sm[threadIdx.x] = make_double2(a, b);
    sm[threadIdx.x + 512] = make_double2(b, a);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < CHAINS; ++k) {
            if (OP == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[k]) : "d"(a), "d"(b));
            if (OP == 1) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[k]) : "d"(a));
            if (OP == 2) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(d[k]) : "d"(a));
            if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[k]) : "f"(fa), "f"(fb));
            if (OP == 4) asm volatile("lg2.approx.f32 %0, %0;" : "+f"(f[k]));
            if (OP == 5) asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f[k]) : "d"(d[k]));
            if (OP == 6) asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d[k]) : "f"(f[k]));
            if (OP == 7) asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(d[k]) : "r"(n[k]));
            if (OP == 8) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(n[k]) : "r"(n[(k + 1) % CHAINS]), "r"(it));
            if (OP == 9) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(n[k]) : "r"(n[(k + 1) % CHAINS]), "r"(it));
            if (OP == 10) {
                int p;
                asm volatile("{.reg .pred q; setp.lt.f64 q, %1, %2; selp.s32 %0, 1, 0, q;}" : "=r"(p) : "d"(d[k]), "d"(a));
                n[k] += p;
            }
            if (OP == 11) {
                double2 v = sm[(threadIdx.x + n[k]) & 1023];
                n[k] += (int)__double2loint(v.x);
            }
            if (OP == 12) asm volatile("min.f64 %0, %0, %1;" : "+d"(d[k]) : "d"(a));
            if (OP == 13) asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(d[k]));
            if (OP == 14) asm volatile("ex2.approx.f32 %0, %0;" : "+f"(f[k]));
            if (OP == 15) asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f[k]) : "r"(n[k]));
            if (OP == 16) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(n[k]) : "r"(n[(k + 1) % CHAINS]), "r"(it));
            if (OP == 17) asm volatile("add.s32 %0, %0, %1;" : "+r"(n[k]) : "r"(it));
        }
    }
    const long long t1 = clock64();
    double acc = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) acc += d[k] + f[k] + n[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char *name, int ctas_per_sm, int threads)
{
    int sms = 148;
    double *out, *in;
    long long *cyc;
    cudaMalloc(&out, sizeof(double) * sms * ctas_per_sm * threads);
    cudaMalloc(&in, sizeof(double) * 64);
    cudaMalloc(&cyc, sizeof(long long) * sms * ctas_per_sm);
    double h[64];
    for (int i = 0; i < 64; ++i) h[i] = 1.0 + i * 1e-3;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    bench<OP><<<sms * ctas_per_sm, threads>>>(out, in, cyc);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<OP><<<sms * ctas_per_sm, threads>>>(out, in, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long hc[148 * 8];
    cudaMemcpy(hc, cyc, sizeof(long long) * sms * ctas_per_sm, cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < sms * ctas_per_sm; ++i) mean += hc[i];
    mean /= sms * ctas_per_sm;
    const double ops_per_sm = (double)ITERS * CHAINS * threads * ctas_per_sm;
    printf("%-22s threads/SM %4d  %8.2f thread-ops/clk/SM  (%.3f ms, %.0f cyc)\n", name, threads * ctas_per_sm,
           ops_per_sm / mean, ms, mean);
    cudaFree(out);
    cudaFree(in);
    cudaFree(cyc);
}

// dependent-chain latency of one op (single warp, one chain per thread)
template <int OP>
__global__ void latency(double *out, const double *in, long long *cycles)
{
    double d = in[threadIdx.x & 63];
    const double a = in[1];
    __shared__ double sm[256];
    sm[threadIdx.x] = a;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 16
    for (int it = 0; it < 4096; ++it) {
        if (OP == 0) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d) : "d"(a));
        if (OP == 1) asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d) : "d"(a));
        if (OP == 2) { d = __dadd_rn(d, sm[(it + threadIdx.x) & 255]); }
    }
    const long long t1 = clock64();
    out[threadIdx.x] = d;
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

template <int OP>
void run_latency(const char *name)
{
    double *out, *in;
    long long *cyc, h;
    cudaMalloc(&out, 8 * 64);
    cudaMalloc(&in, 8 * 64);
    cudaMalloc(&cyc, 8);
    double hin[64];
    for (int i = 0; i < 64; ++i) hin[i] = 1.0 + i * 1e-3;
    cudaMemcpy(in, hin, sizeof hin, cudaMemcpyHostToDevice);
    latency<OP><<<1, 32>>>(out, in, cyc);
    latency<OP><<<1, 32>>>(out, in, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-26s %.2f cycles per dependent op\n", name, (double)h / 4096.0);
}

int main()
{
    run_latency<0>("latency DADD");
    run_latency<1>("latency DFMA");
    run_latency<2>("latency DADD + LDS operand");
    for (int pass = 0; pass < 2; ++pass) {
        const int c = pass == 0 ? 1 : 2, t = 512;
        run<0>("DFMA", c, t);
        run<1>("DADD", c, t);
        run<2>("DMUL", c, t);
        run<12>("DMNMX(min.f64)", c, t);
        run<10>("DSETP+SEL", c, t);
        run<3>("FFMA", c, t);
        run<4>("MUFU.LG2", c, t);
        run<14>("MUFU.EX2", c, t);
        run<13>("MUFU.RCP64H", c, t);
        run<5>("F2F.F32.F64", c, t);
        run<6>("F2F.F64.F32", c, t);
        run<7>("I2F.F64.S32", c, t);
        run<15>("I2F.F32.S32", c, t);
        run<8>("IMAD", c, t);
        run<9>("LOP3", c, t);
        run<16>("SHF", c, t);
        run<17>("IADD", c, t);
        run<11>("LDS.128", c, t);
    }
    return 0;
}
