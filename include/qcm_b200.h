/*
 * qcm_b200.h -- C ABI of the B200 (sm_100a) execution layer behind QCMaquis's contraction::Engine.
 *
 * Plain pointers and sizes only. The C++ host (qcmaquis_b200/csrc/qcm/engine_gpu.hpp, the mirror of
 * contraction::Engine) flattens one (site, direction) contraction into task arrays ONCE, hands them to
 * qcm_plan_create, and then calls qcm_site_hamil2 for every Davidson / Jacobi-Davidson iteration and
 * qcm_boundary_step once per site.  Boundaries live in HBM as qcm_array handles.
 *
 * Conventions follow the reference's own C interface (dmrg/lib/maquis_dmrg/maquis_cinterface.h): global
 * library state, out-arrays of doubles, no callbacks into host code.  Unlike that interface every call
 * returns a status (0 = ok) and qcm_last_error() describes the failure; the C++ Engine mirror turns a
 * non-zero status into std::runtime_error, the reference's error convention on this path
 * (contractions/common/move_boundary.hpp:29, non-abelian/site_hamil.hpp:202).
 *
 * All matrices are column-major FP64 (alps::numeric::matrix<double>, alps/numeric/matrix/matrix.hpp:66).
 */
#ifndef QCM_B200_H
#define QCM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- buffer slots a plan's tasks may reference (Ref.buf) ------------------------------------------- */
enum {
    QCM_BUF_KET_LP = 0,  /* ket, left-paired  [(phys,left),right]   (MPSTensor::make_left_paired,  mpstensor.hpp:155-168) */
    QCM_BUF_KET_RP = 1,  /* ket, right-paired [left,(-phys,right)]  (MPSTensor::make_right_paired, mpstensor.hpp:170-183) */
    QCM_BUF_LEFT   = 2,  /* left boundary,  Boundary::operator[] flattened over b then block (boundary.h:20-145) */
    QCM_BUF_RIGHT  = 3,  /* right boundary */
    QCM_BUF_T      = 4,  /* step-1 products of single-use bonds (BoundaryMPSProduct::at, boundary_times_mps.hpp:163-182) */
    QCM_BUF_TP     = 5,  /* step-1 products of multi-use bonds  (populateData, boundary_times_mps.hpp:188-209) */
    QCM_BUF_Y      = 6,  /* W-applied panels (lbtm/rbtm kernel outputs, abelian/apply_op.hpp, non-abelian/apply_op.hpp) */
    QCM_BUF_OUT    = 7,  /* sigma (left-paired) or the new boundary */
    QCM_BUF_BRA_LP = 8,
    QCM_BUF_BRA_RP = 9,
    QCM_BUF_COUNT  = 10
};

typedef struct { int32_t buf; int32_t pad; int64_t off; } qcm_ref;                       /* element offset inside a slot */

/* panel copy (pairing reshapes: reshapes.h:177-223,289-335; alps_detail.hpp:128-152) */
typedef struct { qcm_ref src, dst; int32_t rows, cols, lds, ldd; } qcm_copy_task;

/* one K-segment of a grouped GEMM: C(+)= alpha * op(A)[m x k] * op(B)[k x n]
 * (block_matrix_algorithms.h:48-162 gemm/gemm_trim_left/right; non-abelian/gemm.hpp:48-204) */
typedef struct { qcm_ref A, B; int32_t lda, ldb, m, n, k, ta, tb, pad; double alpha; } qcm_gemm_seg;
/* one output block = a list of K-segments (the sum over MPO bond terms b that hit the same sector) */
typedef struct { qcm_ref C; int32_t ldc, m, n, seg_begin, seg_end, pad; } qcm_gemm_out;

/* W application (lb_tensor_mpo/rb_tensor_mpo alps_detail.hpp:189-224; SU2 detail::lbtm/rbtm/task_axpy
 * micro_kernels.hpp:19-198) for destination panels that are sums over SEVERAL source panels (a destination with a
 * single source is never materialised: the closing product reads the source panel with the coefficient as alpha).
 * A group holds destination panels fed by (nearly) the same sources: dst_d[e] = sum_u coef[u*ng + d] * src_u[e] over
 * the panel elements e.  coef = W entry * scale * Wigner-9j coupling (gsl_coupling.h:177-204) * Hermitian phase,
 * zero where a destination does not use a source.  All panels of a group are rows x cols, column-major with their own
 * leading dimensions.  cls 1: n_src <= 4, n_dst <= 4, ng = 4 (FMA streaming kernel); cls 0: n_dst <= ng in
 * {8,16,32,64}, coefficient rows padded with zeros to a multiple of 16 sources (DMMA kernel). */
typedef struct { qcm_ref src; int32_t lds, pad; } qcm_w_src;
typedef struct { qcm_ref dst; int32_t ldd, pad; } qcm_w_dst;
typedef struct { int32_t rows, cols, n_src, n_dst, ng, src_begin, dst_begin, cls; int64_t coef_begin; } qcm_w_group;

typedef struct {
    const qcm_gemm_out* t_outs;     int64_t n_t_outs;      /* step 1 for this wave -> QCM_BUF_T   */
    const qcm_gemm_seg* t_segs;     int64_t n_t_segs;
    const qcm_w_group* w_groups;    int64_t n_w_groups;    /* step 2 -> QCM_BUF_Y                 */
    const qcm_w_src* w_srcs;        int64_t n_w_srcs;
    const qcm_w_dst* w_dsts;        int64_t n_w_dsts;
    const double* w_coefs;          int64_t n_w_coefs;
    const qcm_gemm_out* c_outs;     int64_t n_c_outs;      /* step 3 -> QCM_BUF_OUT (accumulating) */
    const qcm_gemm_seg* c_segs;     int64_t n_c_segs;
    int64_t y_elems, t_elems;
    /* exchange wave (sharded plans only, at most one per plan, waves[0]): its W pass writes this rank's partial sums of the
     * destination panels whose sources are spread over several ranks into QCM_BUF_Y[0, world * x_chunk_elems) -- the same
     * layout on every rank; the region is reduce-scattered (rank r receives the complete sums of chunk r) while the
     * remaining waves run, and this wave's closing products (they read chunk `rank`) are executed last.
     * x_zero: the region must be zeroed before the W pass (some panels get no contribution from this rank). */
    int64_t x_chunk_elems;
    int32_t x_zero, pad;
} qcm_wave_desc;

typedef struct {
    int32_t kind;                    /* 0 sigma (site_hamil2), 1 overlap_mpo_left_step, 2 overlap_mpo_right_step, 3 diagonal_hamiltonian */
    int32_t n_waves;
    const qcm_copy_task* pre_copies; int64_t n_pre_copies;
    const qcm_gemm_out* p_outs;      int64_t n_p_outs;     /* multi-use step-1 products -> QCM_BUF_TP */
    const qcm_gemm_seg* p_segs;      int64_t n_p_segs;
    const qcm_wave_desc* waves;
    int64_t elems[QCM_BUF_COUNT];    /* required size (elements) of every slot; inputs are checked, workspaces grown */
    double flops;                    /* algorithmic FLOPs of one execution (schedule-derived) */
    int64_t bytes;                   /* algorithmic bytes of one execution */
    int32_t rank, world;             /* the sharding this plan was built for: world <= 1 means "the whole contraction" (no
                                        collective is issued for it); world > 1 must equal the communicator's size, the
                                        partial results are then summed inside the call */
} qcm_plan_desc;

typedef struct qcm_array_s* qcm_array_t;   /* a device-resident FP64 array */
typedef struct qcm_plan_s*  qcm_plan_t;

/* ---- library state --------------------------------------------------------------------------------- */
int         qcm_init(int device);                 /* selects the device, creates the stream; idempotent */
int         qcm_finalize(void);
const char* qcm_last_error(void);
int         qcm_device_count(int* n);
int         qcm_device_name(char* buf, int len);
int         qcm_sync(void);
void*       qcm_stream(void);                     /* the cudaStream_t every kernel of the library is launched on */
int64_t     qcm_launch_count(void);               /* kernels launched by this library so far */

/* ---- device arrays (HBM-resident boundaries and solver vectors; replaces storage::disk, utils/storage.h) */
int qcm_array_alloc(int64_t n_elems, qcm_array_t* out);
int qcm_array_free(qcm_array_t a);
int qcm_array_size(qcm_array_t a, int64_t* n_elems);
int qcm_array_upload(qcm_array_t a, int64_t off, const double* host, int64_t n);
int qcm_array_download(qcm_array_t a, int64_t off, double* host, int64_t n);
int qcm_array_zero(qcm_array_t a);
void* qcm_array_devptr(qcm_array_t a);
/* ---- spill tier: HBM <-> pinned host memory (the roles of storage::disk::evict / prefetch / fetch, utils/storage.h:113-185,
 * 254-327, with host memory in place of the scratch directory).  qcm_array_evict copies the whole array to `pinned_host`
 * (from qcm_pinned_alloc) on the library's copy stream and releases the device memory, stream-ordered after the work queued so
 * far; the handle stays valid but the array may not be used until qcm_array_prefetch has been called, which allocates device
 * memory again and uploads asynchronously -- work queued later on the library stream waits for the upload, the host never does. */
int qcm_pinned_alloc(int64_t n_elems, double** out);
int qcm_pinned_free(double* p);
int qcm_array_evict(qcm_array_t a, double* pinned_host);
int qcm_array_prefetch(qcm_array_t a, const double* pinned_host);
int qcm_array_resident(qcm_array_t a, int* resident);

/* ---- plans ----------------------------------------------------------------------------------------- */
int qcm_plan_create(const qcm_plan_desc* desc, qcm_plan_t* out);
int qcm_plan_destroy(qcm_plan_t p);
int qcm_plan_stats(qcm_plan_t p, double* flops, int64_t* bytes, int64_t* n_launches, int64_t* workspace_bytes);

/* ---- the three Engine calls -------------------------------------------------------------------------
 * qcm_site_hamil2        replaces contraction::Engine<..>::site_hamil2 (abelian/engine.hpp:196-209,
 *                        non-abelian/engine.hpp:197-207; bodies abelian/site_hamil.hpp:23-90,
 *                        non-abelian/site_hamil.hpp:28-204).  psi/sigma are HOST buffers holding the
 *                        left-paired blocks back to back in DualIndex order; the host<->device copies are
 *                        part of the call.  With a communicator the partial sigma of every rank is summed
 *                        (NCCL allreduce) before the download, so every rank returns the full sigma.
 * qcm_site_hamil2_dev    same on device-resident psi/sigma (solver vectors kept in HBM).
 * qcm_boundary_step      replaces Engine::overlap_mpo_left_step / overlap_mpo_right_step
 *                        (abelian/engine.hpp:102-122; common/move_boundary.hpp:128-229); `in` is the old
 *                        boundary, `out` the new one (device resident, zeroed and filled by the call);
 *                        bra/ket are host buffers (left-paired blocks), consumed before the call returns.  The
 *                        step itself is queued: `out` is complete for every later call of this library and for
 *                        qcm_array_download (stream order); call qcm_sync() to wait for it explicitly. */
int qcm_site_hamil2(qcm_plan_t p, qcm_array_t left, qcm_array_t right, const double* psi, double* sigma);
int qcm_site_hamil2_dev(qcm_plan_t p, qcm_array_t left, qcm_array_t right, qcm_array_t psi, qcm_array_t sigma);
int qcm_boundary_step(qcm_plan_t p, qcm_array_t in, const double* bra, const double* ket, qcm_array_t out);
/* Time-sliced shards: plans[v] = shard v of n of ONE sigma contraction (qcm_plan_sigma(..., rank = v, world = n, ...)), executed
 * one after another on this device, so that only one shard's resident step-1 products (QCM_BUF_TP) occupy HBM at a time -- the
 * single-GPU mode for site problems whose products exceed one device (24e/30o TwoU1 M=4000: 145 GB).  No communicator is
 * involved: the exchange of partial W sums is a local accumulation, sigma accumulates over the shards. */
int qcm_site_hamil2_sliced(const qcm_plan_t* plans, int n, qcm_array_t left, qcm_array_t right, const double* psi, double* sigma);
int qcm_site_hamil2_sliced_dev(const qcm_plan_t* plans, int n, qcm_array_t left, qcm_array_t right, qcm_array_t psi, qcm_array_t sigma);
/* qcm_hdiag              replaces Engine::diagonal_hamiltonian (abelian/engine.hpp:222-227; bodies abelian/h_diag.hpp:41-168,
 *                        non-abelian/h_diag.hpp:19-155): the diagonal of the effective Hamiltonian in the left-paired
 *                        layout of the site tensor (plan kind 3), written to the HOST buffer `diag` (blocks back to back). */
int qcm_hdiag(qcm_plan_t p, qcm_array_t left, qcm_array_t right, double* diag);

/* device time (ms, CUDA events on the library stream) of the kernels of the last plan execution, by phase:
 * [0] reshapes, [1] step-1 GEMMs, [2] W application, [3] closing GEMMs, [4] allreduce, [5] total */
int qcm_last_timing(double ms[6]);
int qcm_set_timing(int enabled);

/* ---- solver-side vector algebra on device arrays (MPSTensor::scalar_overlap / scalar_norm / += / *=,
 *      mpstensor.hpp:346-395,458-522; used by ietl::dot/two_norm, ietl_lanczos_solver.h:67-102) ---------- */
int qcm_vec_dot(qcm_array_t x, qcm_array_t y, int64_t n, double* result);
/* k dot products results[q] = xs[q] . ys[q] with ONE host synchronisation (the Gram-Schmidt / projected-matrix columns of the
 * Jacobi-Davidson driver, ietl/jacobi.h:361-451); the summation order is fixed, so equal inputs give bit-equal results on
 * every rank.  out = sum_j coefs[j] xs[j] in one pass (Ritz vector and residual, jacobi.h:404-420). */
#define QCM_MAX_DOTS 64
int qcm_vec_dots(const qcm_array_t* xs, const qcm_array_t* ys, int k, int64_t n, double* results);
int qcm_vec_lincomb(const qcm_array_t* xs, const double* coefs, int k, qcm_array_t out, int64_t n);
int qcm_vec_axpy(double a, qcm_array_t x, qcm_array_t y, int64_t n);       /* y += a x */
int qcm_vec_scal(double a, qcm_array_t x, int64_t n);
int qcm_vec_copy(qcm_array_t src, qcm_array_t dst, int64_t n);

/* ---- multi-GPU: one process per GPU.  The reference parallelises over the MPO bond index b with OpenMP
 *      (abelian/site_hamil.hpp:74, utils/parallel/loops.hpp:11-26); here the edges (b1, b2) of the MPO bond graph are
 *      sharded over ranks by their step-1 index when the plan is built, every rank executes its plan, and the partial
 *      sigma vectors / boundaries are combined by one allreduce per call inside the library ------------------------- */
int qcm_comm_unique_id(char id[128]);
int qcm_comm_init(int rank, int world, const char id[128]);
int qcm_comm_destroy(void);
int qcm_comm_allreduce(qcm_array_t a, int64_t n);

/* ---- problem descriptors: plan a contraction INSIDE the library from plain arrays ------------------------------------------
 * The entry points above take finished task arrays.  The ones below take the PROBLEM -- the MPO site tensor in CSC form, its
 * operator table as sparse entries, bond spins and Hermitian maps, and the block structures of the site tensor(s) and the
 * boundaries -- and run the schedule builder (qcmaquis_b200/csrc/qcm/plan.hpp) inside libqcm_b200.so.  A QCMaquis-side binding
 * is then a flattening of MPOTensor<Matrix,SymmGroup> (mp_tensors/mpotensor.h:23-107: row/col_dim, the CSC arrays of
 * mpotensor.hpp:10-65, at(b1,b2) -> (tag, scale) lists, left/right_spin, herm_info), of its OPTable entries
 * (block_matrix/site_operator.h: basis(), spin(), get_sparse() -- sparse_operator.h:16-46 (row, col, row_spin, col_spin,
 * coefficient) per block) and of DualIndex / Index lists (dual_index.h:122-338, indexing_stable.hpp); see INTEGRATION.md.
 * Charges are (c0, c1, irrep) as in nu1pg.h:26-84; groups without point group pass irrep = 0. */
typedef struct { int32_t c[3]; } qcm_charge;
typedef struct { qcm_charge q; int64_t size; } qcm_sector;                 /* one entry of an Index<SymmGroup> */
typedef struct { qcm_charge lc, rc; int64_t ls, rs; } qcm_block;           /* one entry of a DualIndex<SymmGroup> */
enum { QCM_SYMM_2U1 = 0, QCM_SYMM_2U1PG = 1, QCM_SYMM_SU2U1 = 2, QCM_SYMM_SU2U1PG = 3 };

typedef struct {
    int32_t spin_twoS, spin_in, spin_out;      /* SpinDescriptor(twoS, in, out), spin_descriptor.h:49-66; zeros for abelian groups */
    int32_t n_blocks;
    const qcm_block* blocks;                   /* op.basis(), DualIndex order */
    const int32_t* entry_ptr;                  /* n_blocks + 1: block b owns entries [entry_ptr[b], entry_ptr[b+1]) */
    const int32_t* row; const int32_t* col;    /* position inside the block */
    const int32_t* row_spin; const int32_t* col_spin;   /* two-site spin labels J, J' of the entry (SU2); NULL: |spin(charge)| */
    const double* coef;
} qcm_site_op_desc;

typedef struct {
    int32_t symm;                              /* QCM_SYMM_* */
    int32_t n_ops;
    const qcm_site_op_desc* ops;               /* the operator table; terms refer to it by index (tag) */
    int64_t row_dim, col_dim, nnz;
    const int64_t* col_ptr;                    /* col_dim + 1 (CSC): entries of column b2, ascending b1 */
    const int64_t* row_idx;                    /* nnz */
    const int64_t* term_ptr;                   /* nnz + 1: entry e owns terms [term_ptr[e], term_ptr[e+1]) in at(b1,b2) order */
    const int32_t* term_op; const double* term_scale;
    const int32_t* left_spin; const int32_t* right_spin;        /* 2S of every bond index (row_dim / col_dim values); NULL for abelian groups */
    const int64_t* left_herm; const int64_t* right_herm;        /* Hermitian::LeftHerm / RightHerm (mpotensor_detail.h:118-168); NULL: no pairs */
    const int32_t* left_phase; const int32_t* right_phase;
} qcm_mpo_desc;

typedef struct {                               /* block structure of one MPSTensor; its data travel as left-paired blocks back to back */
    int32_t n_phys, n_left, n_right, n_blocks;
    const qcm_sector* phys; const qcm_sector* left; const qcm_sector* right;   /* site_dim(), row_dim(), col_dim() */
    const qcm_block* blocks;                   /* data().basis() after make_left_paired() */
} qcm_tensor_desc;

typedef struct {                               /* block structure of a Boundary: entry b owns blocks [block_ptr[b], block_ptr[b+1]) */
    int64_t aux_dim;
    const int64_t* block_ptr;
    const qcm_block* blocks;
} qcm_boundary_desc;

typedef struct qcm_mpo_s* qcm_mpo_t;
int qcm_mpo_upload(const qcm_mpo_desc* d, qcm_mpo_t* out);      /* once per MPO site tensor (or fused two-site tensor) */
int qcm_mpo_free(qcm_mpo_t m);
/* rank/world: the shard to plan (0, 1: the whole contraction); ws_budget_elems: workspace budget in elements (0: default, 2^32) */
int qcm_plan_sigma(qcm_mpo_t m, const qcm_tensor_desc* ket, const qcm_boundary_desc* left, const qcm_boundary_desc* right,
                   int rank, int world, int64_t ws_budget_elems, qcm_plan_t* out);
int qcm_plan_left_step(qcm_mpo_t m, const qcm_tensor_desc* bra, const qcm_tensor_desc* ket, const qcm_boundary_desc* left,
                       int rank, int world, int64_t ws_budget_elems, qcm_plan_t* out);
int qcm_plan_right_step(qcm_mpo_t m, const qcm_tensor_desc* bra, const qcm_tensor_desc* ket, const qcm_boundary_desc* right,
                        int rank, int world, int64_t ws_budget_elems, qcm_plan_t* out);
/* Noise term of the perturbed density matrix -- replaces Engine::left_boundary_tensor_mpo / right_boundary_tensor_mpo
 * (dmrg/framework/dmrg/mp_tensors/contractions/common/move_boundary.hpp:68-126) TOGETHER WITH the accumulation loops of their only
 * callers (contractions/common/prediction.hpp:34-47,101-114; mp_tensors/twositetensor.hpp:192-219,260-287):
 *   left:  sum over b2 of Y[b2] Y[b2]^T,   right: sum over b1 of Y'[b1]^T Y'[b1]   (all blocks; the caller keeps those its dm has).
 * Execute with qcm_boundary_step(plan, boundary, ket, ket, out); out holds ONE bond entry (the density-matrix blocks), described
 * by qcm_plan_out_size / qcm_plan_out_blocks.  The term is quadratic in the half-contracted boundary and is not sharded. */
int qcm_plan_noise_left(qcm_mpo_t m, const qcm_tensor_desc* ket, const qcm_boundary_desc* left, int64_t ws_budget_elems, qcm_plan_t* out);
int qcm_plan_noise_right(qcm_mpo_t m, const qcm_tensor_desc* ket, const qcm_boundary_desc* right, int64_t ws_budget_elems, qcm_plan_t* out);
/* block structure of the result of a plan made by the calls above: sigma (aux_dim 1) or the new boundary.
 * qcm_plan_out_size: number of bond entries, blocks and elements; qcm_plan_out_blocks: block_ptr (aux_dim + 1), the blocks in
 * DualIndex order and the element offset of every block inside the output array. */
int qcm_plan_out_size(qcm_plan_t p, int64_t* aux_dim, int64_t* n_blocks, int64_t* n_elems);
int qcm_plan_out_blocks(qcm_plan_t p, int64_t* block_ptr, qcm_block* blocks, int64_t* elem_off);

/* ---- measured FP64 peaks (roofline denominators; not part of the contraction path) ------------------ */
int qcm_measure_fp64_fma_peak(double* tflops);     /* register-resident DFMA chains on all SMs */
int qcm_measure_fp64_dmma_peak(double* tflops);    /* register-resident mma.sync m8n8k4 f64 chains */
int qcm_measure_hbm_copy(double* gbs);

#ifdef __cplusplus
}
#endif
#endif
