// qcm_dev.cuh -- device-side task records shared by the translation units of libqcm_b200.so
#pragma once
#include "../../include/qcm_b200.h"
#include <cuda_runtime.h>

struct BufTable { double* p[QCM_BUF_COUNT]; };

// one K-segment of a grouped GEMM and one output tile ("work item") with its K-segment range
struct DSeg { long long a_off, b_off; int a_buf, b_buf, lda, ldb, m, n, k, ta, tb, pad; double alpha; };
struct DWork { long long c_off; int c_buf, ldc, m0, n0, m, n, seg_begin, seg_end, mode, pad; };   // mode 0 store, 1 add, 2 atomic

// grouped GEMM, persistent warp-specialised kernels (gemm_ws.cu)
struct GemmWsVariant { int tm, tn, threads; double eff; };
int gemm_ws_num_variants();
GemmWsVariant gemm_ws_variant(int v);
const char* gemm_ws_init(int sm_count);     // sets kernel attributes, queries occupancy; returns nullptr or an error text
int gemm_ws_grid(int v, long long n_works); // CTAs to launch for n_works work items
void gemm_ws_launch(int v, long long n_works, const DWork* works, const DSeg* segs, BufTable const& bufs, cudaStream_t st);
