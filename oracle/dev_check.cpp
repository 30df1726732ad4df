// development driver (not part of the test-suite): quick look at goldens with the oracle engine
#include "oracle_engine.hpp"
#include "qcm/scenarios.hpp"
#include <cstdio>
#include <cstring>

using namespace qcm;

static void report_mpo(Problem& P)
{
    printf("terms=%zu  bonds:", P.model->terms.size());
    for (size_t p = 0; p < P.mpo.size(); ++p) printf(" %zu/%zu", P.mpo[p].col_dim(), P.mpo.herm_pairs[p]);
    printf("  core=%.12f\n", P.mpo.core_energy);
}

int main(int argc, char** argv)
{
    std::string which = argc > 1 ? argv[1] : "h2";
    std::string symm = argc > 2 ? argv[2] : "su2u1";
    Problem P;
    P.params.symm = symm_from_string(symm);
    if (which == "h2") {
        P.params.L = 2; P.params.site_types = {0, 0};
        double v[] = {0.354237848011, -0.821703816101E-13, 0.185125251547, 0.782984788117E-13, 0.361001163519, 0.371320200119,
                      -0.678487901790, -0.539801158857E-14, -0.653221638776, 0.176392403557};
        int ix[][4] = {{1,1,1,1},{1,1,2,1},{2,1,2,1},{2,2,2,1},{1,1,2,2},{2,2,2,2},{1,1,0,0},{2,1,0,0},{2,2,0,0},{0,0,0,0}};
        for (int i = 0; i < 10; ++i) P.params.integrals.push_back(Integral{{ix[i][0], ix[i][1], ix[i][2], ix[i][3]}, v[i]});
        P.params.integral_cutoff = 1e-100;   // default integral_cutoff? (test uses default)
        P.params.nelec = 2; P.params.spin = 0; P.params.nup = 1; P.params.ndown = 1;
    } else {
        P.params.integrals = read_fcidump(which);
        int L = atoi(argv[3]);
        P.params.L = L; P.params.site_types.assign(L, 0);
        P.params.nelec = atoi(argv[4]); P.params.spin = 0; P.params.nup = P.params.nelec / 2; P.params.ndown = P.params.nelec / 2;
    }
    P.build_model();
    P.build_mpo();
    report_mpo(P);
    oracle::OracleEngine eng(P.params.symm);
    int L = P.params.L;
    // random MPS, full chain energy identity
    P.init_mps(100, true, 0., 42);
    P.build_boundaries(eng);
    double e_left = P.left[L][0].trace();
    double e_right = P.right[0][0].trace();
    printf("chain <H> from left = %.14f, from right = %.14f (blocks %zu, aux %zu)\n", e_left, e_right, P.left[L][0].n_blocks(), P.left[L].aux_dim());
    for (int p = 0; p < L; ++p) {
        MPSTensor s = eng.site_hamil2(P.mps[p], P.left[p], P.right[p + 1], P.mpo[p]);
        printf("  site %d  <psi|sigma> = %.14f\n", p, s.scalar_overlap(P.mps[p]));
    }
    // exact: edges = complete 1-dim sectors, two-site centre
    if (L == 2 || L == 4) {
        std::vector<Index> allowed = allowed_sectors(P.params.symm, P.site_types(), P.model->phys_indices, P.model->total_charge, 1000);
        MPS ex;
        for (int p = 0; p < L; ++p) ex.push_back(MPSTensor(P.phys(p), allowed[p], allowed[p + 1], []() { return 1.0; }));
        P.mps = ex;
        int c = L / 2 - 1;
        P.build_boundaries(eng, c, c + 2);
        MPOTensor const& ts = P.twosite_mpo(c);
        printf("two-site MPO %zux%zu nnz %zu\n", ts.row_dim(), ts.col_dim(), ts.nnz());
        MPSTensor templ = make_twosite_tensor(P.phys(c), P.phys(c + 1), allowed[c], allowed[c + 2], []() { return 1.0; });
        double asym = 0;
        std::vector<double> w = dense_heff_spectrum(eng, templ, P.left[c], P.right[c + 2], ts, &asym);
        printf("dense Heff n=%zu asym=%.3e  E0 = %.14f  (E0+core = %.14f)\n", w.size(), asym, w[0], w[0] + P.mpo.core_energy);
    }
    return 0;
}
