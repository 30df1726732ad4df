// development tool: statistics of a sigma plan (fan-in / fan-out of the W application, GEMM shape histograms)
#include "qcm/scenarios.hpp"
#include "qcm/plan.hpp"
#include <cstdio>
using namespace qcm;
int main(int argc, char** argv)
{
    if (argc < 7) { printf("usage: plan_stats fcidump symm L nelec site M\n"); return 1; }
    Problem P; P.params.symm = symm_from_string(argv[2]); P.params.integrals = read_fcidump(argv[1]);
    int L = atoi(argv[3]); P.params.L = L; P.params.site_types.assign(L, 0);
    P.params.nelec = atoi(argv[4]); P.params.spin = 0; P.params.nup = P.params.nelec / 2; P.params.ndown = P.params.nelec - P.params.nelec / 2;
    P.build_model(); P.build_mpo();
    int site = atoi(argv[5]); size_t M = atoi(argv[6]);
    SyntheticSite S = make_synthetic_site(P, site, true, M, 1);
    S.psi.make_left_paired();
    std::vector<DualIndex> lb(S.left.aux_dim()), rb(S.right.aux_dim());
    for (size_t k = 0; k < lb.size(); ++k) lb[k] = S.left[k].basis();
    for (size_t k = 0; k < rb.size(); ++k) rb[k] = S.right[k].basis();
    plan::BoundaryLayout ll, rl; ll.assign(lb); rl.assign(rb);
    plan::TensorDesc td{S.psi.site_dim(), S.psi.row_dim(), S.psi.col_dim(), S.psi.data().basis()};
    if (argc > 7) {     // sharding quality: per-rank FLOPs for world sizes 1, 2, 4, 8
        for (int world : {1, 2, 4, 8}) {
            double mx = 0, sum = 0, t_sum = 0, alg = 0;
            for (int r = 0; r < world; ++r) {
                plan::Planner plr(P.symm(), *S.mpo, true, r, world, (int64_t)1 << 40);
                plan::Plan q = plr.plan_sigma(td, ll, rl);
                double f = q.flops_t + q.exec_w + q.exec_close;
                mx = std::max(mx, f); sum += f; t_sum += q.flops_t; alg += q.flops();
                printf("  world %d rank %d: step1 %.3e  W %.3e  close %.3e  total %.3e  TP %.2f GB  exchange region %.2f GB (chunk %.3f GB = %lld elements)\n", world, r, q.flops_t, q.exec_w, q.exec_close, f,
                       q.tp_elems * 8e-9, (!q.waves.empty() ? q.waves[0].x_chunk * world * 8e-9 : 0.), (!q.waves.empty() ? q.waves[0].x_chunk * 8e-9 : 0.), (long long)(!q.waves.empty() ? q.waves[0].x_chunk : 0));
            }
            printf("world %d: max rank FLOPs %.3e, sum %.3e (step 1 %.3e); algorithmic FLOPs booked over ranks %.10e\n", world, mx, sum, t_sum, alg);
        }
        return 0;
    }
    plan::Planner pl(P.symm(), *S.mpo, true, 0, 1, (int64_t)1 << 40);
    plan::Plan pp = pl.plan_sigma(td, ll, rl);
    {   // digest of everything the device executes (to compare planner versions / thread counts)
        uint64_t h = 1469598103934665603ull;
        auto mix = [&](const void* p, size_t n) { const unsigned char* c = (const unsigned char*)p; for (size_t i = 0; i < n; ++i) { h ^= c[i]; h *= 1099511628211ull; } };
        auto gl = [&](plan::GemmList const& g) {
            for (auto const& o : g.outs) { mix(&o.C.buf, 4); mix(&o.C.off, 8); mix(&o.ldc, 4); mix(&o.m, 4); mix(&o.n, 4); mix(&o.seg_begin, 4); mix(&o.seg_end, 4); }
            for (auto const& s : g.segs) { mix(&s.A.buf, 4); mix(&s.A.off, 8); mix(&s.B.buf, 4); mix(&s.B.off, 8); mix(&s.lda, 4); mix(&s.ldb, 4); mix(&s.m, 4); mix(&s.n, 4); mix(&s.k, 4); mix(&s.ta, 4); mix(&s.tb, 4); mix(&s.alpha, 8); } };
        gl(pp.persistent_t);
        for (auto const& W : pp.waves) {
            gl(W.t_gemm); gl(W.close_gemm);
            for (auto const& g : W.w_groups.groups) { mix(&g.rows, 4); mix(&g.cols, 4); mix(&g.n_src, 4); mix(&g.n_dst, 4); mix(&g.ng, 4); mix(&g.cls, 4); mix(&g.src_begin, 4); mix(&g.dst_begin, 4); mix(&g.coef_begin, 8); }
            for (auto const& s : W.w_groups.srcs) { mix(&s.src.buf, 4); mix(&s.src.off, 8); mix(&s.lds, 4); }
            for (auto const& d : W.w_groups.dsts) { mix(&d.dst.buf, 4); mix(&d.dst.off, 8); mix(&d.ldd, 4); }
            mix(W.w_groups.coefs.data(), W.w_groups.coefs.size() * 8);
        }
        printf("plan digest %016llx\n", (unsigned long long)h);
    }
    printf("waves %zu  flops t %.3e w %.3e close %.3e\n", pp.waves.size(), pp.flops_t, pp.flops_w, pp.flops_close);
    printf("elems: left %.3e right %.3e psi %.3e  TP %.3e  T %.3e  Y %.3e\n", (double)ll.total, (double)rl.total, (double)pp.ket_lp_elems, (double)pp.tp_elems, (double)pp.t_elems_max, (double)pp.y_elems_max);
    printf("W: groups %lld elems read %.3e written %.3e\n", (long long)pp.w_groups, (double)pp.w_elems_read, (double)pp.w_elems_written);
    printf("panels: direct %.3e  via W %.3e  skipped (no closing product) %.3e elements;  executed flops: W %.3e close %.3e\n", (double)pp.direct_panel_elems, (double)pp.w_panel_elems, (double)pp.skipped_panel_elems, pp.exec_w, pp.exec_close);
    { double rd[2] = {0, 0}, wr[2] = {0, 0}, padf = 0; long long ng[2] = {0, 0}; for (auto const& W : pp.waves) for (auto const& g : W.w_groups.groups) { double el = (double)g.rows * g.cols; rd[g.cls] += el * g.n_src; wr[g.cls] += el * g.n_dst; ng[g.cls]++; if (!g.cls) padf += 2.0 * el * ((g.n_src + 7) / 8 * 8) * ((g.n_dst + 7) / 8 * 8); }
      printf("   gemm class: %lld groups read %.3e written %.3e (DMMA flops incl. padding %.3e);  stream class: %lld groups read %.3e written %.3e\n", ng[0], rd[0], wr[0], padf, ng[1], rd[1], wr[1]); }
    // fan-in histogram weighted by panel elements
    double fin[8] = {0}, tot = 0, pair_el = 0; const int edges[8] = {1, 2, 4, 8, 16, 64, 256, 1 << 30};
    double fout_src_elems = 0; std::map<std::pair<int, int64_t>, int> fanout; std::map<std::pair<int, int64_t>, double> src_el;
    double gsz[65] = {0};
    for (auto const& W : pp.waves) {
        auto const& wl = W.w_groups;
        for (auto const& g : wl.groups) {
            double el = (double)g.rows * g.cols;
            gsz[g.n_dst] += el * g.n_dst;
            for (int d = 0; d < g.n_dst; ++d) {
                int n = 0;
                for (int u = 0; u < g.n_src; ++u) if (wl.coefs[g.coef_begin + (size_t)u * g.ng + d] != 0.) { ++n; auto key = std::make_pair(wl.srcs[g.src_begin + u].src.buf, wl.srcs[g.src_begin + u].src.off); fanout[key]++; src_el[key] = el; }
                for (int e = 0; e < 8; ++e) if (n <= edges[e]) { fin[e] += el; break; }
                tot += el; pair_el += el * n;
            }
        }
    }
    printf("dst panel elements %.3e, (src,dst) pair elements %.3e (avg fan-in %.2f)\n", tot, pair_el, pair_el / tot);
    printf("fan-in (sources per destination panel), share of destination elements:\n");
    for (int e = 0; e < 8; ++e) printf("  <=%-10d %6.2f %%\n", edges[e], 100 * fin[e] / tot);
    double fo[8] = {0}, stot = 0;
    for (auto const& kv : fanout) { double el = src_el[kv.first]; stot += el; for (int e = 0; e < 8; ++e) if (kv.second <= edges[e]) { fo[e] += el; break; } }
    printf("distinct source panel elements %.3e; fan-out (destinations per source panel), share of source elements:\n", stot);
    for (int e = 0; e < 8; ++e) printf("  <=%-10d %6.2f %%\n", edges[e], 100 * fo[e] / stot);
    printf("group size (destinations per group), share of destination elements:\n");
    for (int i = 1; i <= 64; ++i) if (gsz[i] > 0) printf("  %2d %6.2f %%\n", i, 100 * gsz[i] / tot);
    // GEMM shape histograms (flop weighted)
    auto shape = [&](const char* name, std::vector<plan::GemmList const*> lists) {
        double byk[6] = {0}, bym[6] = {0}, byn[6] = {0}, f = 0; const int ed[6] = {8, 16, 32, 64, 128, 1 << 30}; size_t nouts = 0, nsegs = 0;
        for (auto gl : lists) { nouts += gl->outs.size(); nsegs += gl->segs.size();
            for (auto const& s : gl->segs) { double fl = 2.0 * s.m * s.n * s.k; f += fl;
                for (int e = 0; e < 6; ++e) if (s.k <= ed[e]) { byk[e] += fl; break; }
                for (int e = 0; e < 6; ++e) if (s.m <= ed[e]) { bym[e] += fl; break; }
                for (int e = 0; e < 6; ++e) if (s.n <= ed[e]) { byn[e] += fl; break; } } }
        printf("%s: %zu outputs, %zu segments, %.3e flops; flop share by dimension bucket (<=8,16,32,64,128,more)\n", name, nouts, nsegs, f);
        printf("   m:"); for (int e = 0; e < 6; ++e) printf(" %5.1f", 100 * bym[e] / f); printf("\n   n:"); for (int e = 0; e < 6; ++e) printf(" %5.1f", 100 * byn[e] / f);
        printf("\n   k:"); for (int e = 0; e < 6; ++e) printf(" %5.1f", 100 * byk[e] / f); printf("\n");
    };
    std::vector<plan::GemmList const*> tl{&pp.persistent_t}, cl;
    for (auto const& W : pp.waves) { tl.push_back(&W.t_gemm); cl.push_back(&W.close_gemm); }
    shape("step 1", tl); shape("step 3", cl);
    return 0;
}
