// qcm_dev.cuh -- device-side task records shared by the translation units of libqcm_b200.so
#pragma once
#include "../../include/qcm_b200.h"
#include <cuda_runtime.h>

struct BufTable { double* p[QCM_BUF_COUNT]; };

// one K-segment of a grouped GEMM and one output tile ("work item") with its K-segment range
struct DSeg { long long a_off, b_off; int a_buf, b_buf, lda, ldb, m, n, k, ta, tb, pad; double alpha; };
struct DWork { long long c_off; int c_buf, ldc, m0, n0, m, n, seg_begin, seg_end, mode, pad; };   // mode 0 store, 1 add, 2 atomic

// W application (gathered dense products): source / destination panel references, groups and work items
struct DWSrc { long long off; int buf, lds; };
struct DWDst { long long off; int buf, ldd; };
struct DWGroup { int rows, cols, n_src, n_dst, ng, src_begin, dst_begin, cls; long long coef_begin; };
struct DWWork { int group, e0; };   // panel elements [e0, e0 + tile) of a group (element e = row + col * rows)

#ifdef __CUDACC__
// ---- device helpers shared by the warp-specialised kernels ------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b)
{
    asm volatile("{\n .reg .b64 t;\n mbarrier.arrive.shared::cta.b64 t, [%0];\n}\n" ::"r"(smem_u32(b)) : "memory");
}
// arrives on b (without incrementing its pending count) once all cp.async issued so far by this thread have landed
__device__ __forceinline__ void mbar_cp_async_arrive(unsigned long long* b)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity)
{
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        "QCM_WAIT:\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        " @p bra QCM_DONE;\n"
        " bra QCM_WAIT;\n"
        "QCM_DONE:\n"
        "}\n" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc, bool valid)
{
    const int bytes = valid ? 8 : 0;    // src-size 0: nothing is read, the 8 destination bytes are zero-filled
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes) : "memory");
}
// same with a precomputed shared-memory address and byte count (0 or 8)
__device__ __forceinline__ void cp_async8_s(unsigned smem_addr, const void* gsrc, int bytes)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(smem_addr), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void dmma8x8x4(double& c0, double& c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
#endif

// W application, persistent warp-specialised kernels (wgemm_ws.cu); class c = 0..3 for ng = 8, 16, 32, 64
const char* wgemm_ws_init(int sm_count);
int wgemm_ws_tile();                        // panel elements per work item
void wgemm_ws_launch(int c, long long n_works, const DWWork* works, const DWGroup* groups, const DWSrc* srcs, const DWDst* dsts, const double* coefs,
                     BufTable const& bufs, cudaStream_t st);

// grouped GEMM, persistent warp-specialised kernels (gemm_ws.cu)
struct GemmWsVariant { int tm, tn, threads; double eff; };
int gemm_ws_num_variants();
GemmWsVariant gemm_ws_variant(int v);
const char* gemm_ws_init(int sm_count);     // sets kernel attributes, queries occupancy; returns nullptr or an error text
int gemm_ws_grid(int v, long long n_works); // CTAs to launch for n_works work items
void gemm_ws_launch(int v, long long n_works, const DWork* works, const DSeg* segs, BufTable const& bufs, cudaStream_t st);
