// TEST / MEASUREMENT INFRASTRUCTURE -- C entry points of the CPU oracle (ctypes).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg load this library.
// It rebuilds the same synthetic site problem as the GPU driver (same generator, same seeds) and times the
// oracle's site_hamil2 on the host cores: OpenMP over the MPO bond index with one BLAS thread per task, the
// reference's own threading model (utils/parallel/loops.hpp:11-26).
#include "oracle_engine.hpp"
#include "qcm/scenarios.hpp"
#include "qcm/sweep.hpp"
#include "qcm/twosite.hpp"
#include <cstdio>
#include <cstring>
#include <omp.h>

using namespace qcm;
extern "C" void scipy_openblas_set_num_threads(int);

namespace {
struct Orc { Problem P; SyntheticSite S; MPSTensor sigma; };
void set_err(char* err, int errlen, std::string const& s) { if (err && errlen > 0) snprintf(err, errlen, "%s", s.c_str()); }
}

extern "C" void* orc_create(const char* fcidump, const char* symm, int L, int nelec, char* err, int errlen)
{
    try {
        std::unique_ptr<Orc> D(new Orc());
        Problem& P = D->P;
        P.params.symm = symm_from_string(symm);
        P.params.integrals = read_fcidump(fcidump);
        P.params.L = L; P.params.site_types.assign(L, 0);
        P.params.nelec = nelec; P.params.spin = 0; P.params.nup = nelec / 2; P.params.ndown = nelec - nelec / 2;
        P.build_model();
        P.build_mpo();
        return D.release();
    } catch (std::exception const& e) { set_err(err, errlen, e.what()); return nullptr; }
}
extern "C" void orc_destroy(void* h) { delete static_cast<Orc*>(h); }
extern "C" int orc_threads() { return omp_get_max_threads(); }
// torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm uses all host cores like the reference's OpenMP build
extern "C" void orc_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }

extern "C" int orc_setup_site(void* h, int site, int twosite, int M, unsigned seed, double* psi_elems, char* err, int errlen)
{
    try {
        Orc* D = static_cast<Orc*>(h);
        D->S = make_synthetic_site(D->P, site, twosite != 0, (size_t)M, seed);
        D->S.psi.make_left_paired();
        *psi_elems = (double)D->S.psi.data().num_elements();
        return 0;
    } catch (std::exception const& e) { set_err(err, errlen, e.what()); return 1; }
}

// n sigma evaluations; seconds = wall time of all of them; overlap = <psi|sigma>
extern "C" int orc_sigma(void* h, int n, double* seconds, double* overlap, double* sigma_elems, char* err, int errlen)
{
    try {
        Orc* D = static_cast<Orc*>(h);
        scipy_openblas_set_num_threads(1);
        oracle::OracleEngine eng(D->P.symm());
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < n; ++i) D->sigma = eng.site_hamil2(D->S.psi, D->S.left, D->S.right, *D->S.mpo);
        *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        D->sigma.make_left_paired();
        *overlap = D->sigma.scalar_overlap(D->S.psi);
        *sigma_elems = (double)D->sigma.data().num_elements();
        return 0;
    } catch (std::exception const& e) { set_err(err, errlen, e.what()); return 1; }
}
// left-paired blocks of the last sigma, back to back in DualIndex order
extern "C" int orc_get_sigma(void* h, double* out)
{
    Orc* D = static_cast<Orc*>(h);
    size_t o = 0;
    for (size_t k = 0; k < D->sigma.data().n_blocks(); ++k) { auto const& v = D->sigma.data()[k].v; std::memcpy(out + o, v.data(), v.size() * 8); o += v.size(); }
    return 0;
}

// the same single-site sweeps (qcm/sweep.hpp) on the CPU oracle: the sweep-level CPU baseline and energy reference
extern "C" int orc_ss_sweeps(void* h, int Mmax, int nsweeps, unsigned seed, double* energies, int n_max, int* n_out, double* info, char* err, int errlen)
{
    try {
        Orc* D = static_cast<Orc*>(h);
        scipy_openblas_set_num_threads(1);
        D->P.init_mps((size_t)Mmax, true, 0., seed);
        oracle::OracleEngine eng(D->P.symm());
        sweep::SweepLog log = sweep::ss_sweeps(eng, D->P.mpo, D->P.mps, nsweeps);
        int n = (int)std::min<size_t>(log.energies.size(), (size_t)n_max);
        for (int i = 0; i < n; ++i) energies[i] = log.energies[i];
        *n_out = n;
        double secs = 0; for (double s : log.sweep_seconds) secs += s;
        info[0] = (double)log.total_sigma; info[1] = secs; info[2] = log.energies.back();
        return 0;
    } catch (std::exception const& e) { set_err(err, errlen, e.what()); return 1; }
}

// the same two-site sweeps (qcm/twosite.hpp) on the CPU oracle
extern "C" int orc_ts_sweeps(void* h, int M0, int Mmax, int nsweeps, unsigned seed, double* energies, int n_max, int* n_out, double* info, char* err, int errlen)
{
    try {
        Orc* D = static_cast<Orc*>(h);
        scipy_openblas_set_num_threads(1);
        D->P.init_mps((size_t)M0, true, 0., seed);
        oracle::OracleEngine eng(D->P.symm());
        ts::TsParams prm; prm.Mmax = (size_t)Mmax;
        std::vector<size_t> dims;
        sweep::SweepLog log = ts::ts_sweeps(D->P.symm(), eng, D->P.mpo, [&](int p) -> MPOTensor const& { return D->P.twosite_mpo(p); }, D->P.mps, nsweeps, prm, &dims);
        int n = (int)std::min<size_t>(log.energies.size(), (size_t)n_max);
        for (int i = 0; i < n; ++i) energies[i] = log.energies[i];
        *n_out = n;
        double secs = 0; for (double s : log.sweep_seconds) secs += s;
        info[0] = (double)log.total_sigma; info[1] = secs; info[2] = log.energies.back();
        info[3] = dims.empty() ? 0. : (double)*std::max_element(dims.begin(), dims.end());
        return 0;
    } catch (std::exception const& e) { set_err(err, errlen, e.what()); return 1; }
}

// the synthetic-start two-site sweeps of qcmd_ts_sweeps_synth (driver_capi.cpp) on the CPU oracle: same starting state
// (make_synthetic_mps), same driver, same parameters
extern "C" int orc_ts_sweeps_synth(void* h, int M, int nsweeps, unsigned seed, int max_micro, double* energies, int n_max, int* n_out, double* info, char* err, int errlen)
{
    try {
        Orc* D = static_cast<Orc*>(h);
        scipy_openblas_set_num_threads(1);
        D->P.mps = make_synthetic_mps(D->P, (size_t)M, seed);
        oracle::OracleEngine eng(D->P.symm());
        ts::TsParams prm; prm.Mmax = (size_t)M; prm.drop_stale = true; prm.max_micro_iterations = max_micro;
        std::vector<size_t> dims;
        sweep::SweepLog log = ts::ts_sweeps(D->P.symm(), eng, D->P.mpo, [&](int p) -> MPOTensor const& { return D->P.twosite_mpo(p); }, D->P.mps, nsweeps, prm, &dims);
        int n = (int)std::min<size_t>(log.energies.size(), (size_t)n_max);
        for (int i = 0; i < n; ++i) energies[i] = log.energies[i];
        *n_out = n;
        double secs = 0; for (double s : log.sweep_seconds) secs += s;
        info[0] = (double)log.total_sigma; info[1] = secs; info[2] = log.energies.empty() ? 0. : log.energies.back();
        info[3] = dims.empty() ? 0. : (double)*std::max_element(dims.begin(), dims.end());
        return 0;
    } catch (std::exception const& e) { set_err(err, errlen, e.what()); return 1; }
}
