// FP64 pipe probes for sm_100a: DFMA chains and the mma.sync f64 shapes (m8n8k4, m16n8k4, m16n8k8, m16n8k16),
// register-resident, at several warps per SM.  Prints TFLOP/s per variant.  Build: nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE> __device__ __forceinline__ void mma(double (&c)[4], const double (&a)[8], const double (&b)[4])
{
    if (SHAPE == 0)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c[0]), "+d"(c[1]) : "d"(a[0]), "d"(b[0]));
    else if (SHAPE == 1)
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                     : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
    else if (SHAPE == 2)
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                     : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                     : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int SHAPE, int NACC> __global__ void k_mma(double* out, int iters)
{
    double c[NACC][4], a[8], b[4];
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
    for (int i = 0; i < 4; ++i) b[i] = 1e-3 * (i + 1);
    for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) c[i][j] = i + j;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < NACC; ++i) mma<SHAPE>(c[i], a, b);
    double s = 0;
    for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    if (s == 123.456) out[0] = s;
}
template <int NACC> __global__ void k_fma(double* out, int iters)
{
    double a[NACC], x = 1.0000001 + threadIdx.x * 1e-9, y = 1e-9;
    for (int i = 0; i < NACC; ++i) a[i] = i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < NACC; ++i) a[i] = fma(a[i], x, y);
    double s = 0;
    for (int i = 0; i < NACC; ++i) s += a[i];
    if (s == 123.456) out[0] = s;
}
// DFMA and DMMA issued from different warps of the same SM: do the two paths add up?
template <int NACC> __global__ void k_mix(double* out, int iters)
{
    if ((threadIdx.x >> 5) & 1) {
        double c[NACC][4], a[8], b[4];
        for (int i = 0; i < 8; ++i) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
        for (int i = 0; i < 4; ++i) b[i] = 1e-3 * (i + 1);
        for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) c[i][j] = i + j;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < NACC; ++i) mma<0>(c[i], a, b);
        double s = 0;
        for (int i = 0; i < NACC; ++i) for (int j = 0; j < 2; ++j) s += c[i][j];
        if (s == 123.456) out[0] = s;
    } else {
        double a[NACC], x = 1.0000001 + threadIdx.x * 1e-9, y = 1e-9;
        for (int i = 0; i < NACC; ++i) a[i] = i;
        for (int it = 0; it < iters * 8; ++it)     // 8 DFMA per DMMA slot: 32 lanes x 8 = 256 FMA = one m8n8k4
#pragma unroll
            for (int i = 0; i < NACC; ++i) a[i] = fma(a[i], x, y);
        double s = 0;
        for (int i = 0; i < NACC; ++i) s += a[i];
        if (s == 123.456) out[0] = s;
    }
}

template <class F> float time_ms(F f)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best; }
    return best;
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    printf("device %s, %d SMs, clock %d kHz\n", p.name, sms, p.clockRate);
    double* out; cudaMalloc(&out, 64);
    const int iters = 4000;
    const double fl[4] = {2.0 * 8 * 8 * 4, 2.0 * 16 * 8 * 4, 2.0 * 16 * 8 * 8, 2.0 * 16 * 8 * 16};
    const char* nm[4] = {"m8n8k4", "m16n8k4", "m16n8k8", "m16n8k16"};
    for (int wps : {4, 8, 16, 32}) {          // warps per SM (one CTA per SM)
        int threads = wps * 32;
        float ms;
        ms = time_ms([&] { k_fma<8><<<sms, threads>>>(out, iters * 8); });
        printf("warps/SM %2d  DFMA x8acc      %7.2f TF/s\n", wps, 2.0 * 8 * iters * 8 * (double)sms * threads / (ms * 1e-3) / 1e12);
#define RUN(S, N) ms = time_ms([&] { k_mma<S, N><<<sms, threads>>>(out, iters); }); \
        printf("warps/SM %2d  %-9s x%dacc %7.2f TF/s\n", wps, nm[S], N, fl[S] * N * iters * (double)sms * wps / (ms * 1e-3) / 1e12);
        RUN(0, 4) RUN(0, 8) RUN(0, 16) RUN(1, 4) RUN(1, 8) RUN(2, 4) RUN(2, 8) RUN(3, 2) RUN(3, 4) RUN(3, 8)
        ms = time_ms([&] { k_mix<8><<<sms, threads>>>(out, iters); });
        printf("warps/SM %2d  DFMA+DMMA mixed %7.2f TF/s\n", wps, (fl[0] * 8 * iters * (double)sms * wps / 2 + 2.0 * 8 * iters * 8 * (double)sms * threads / 2) / (ms * 1e-3) / 1e12);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
