// ell_views.cpp -- CPU check of the index views the two-slot kernel (osqp-solver_b200/csrc/gomp_duo.cu) walks: the
// "ELL by owner" views of A and P (symbolic.h: build_owner_ell) and the assembly lists of S = P + sigma I + A' diag(rho) A
// (build_assembly_ell), against plain dense / CSC algebra on random values.  Test infrastructure (no GPU): a wrong entry in
// one of these tables would otherwise only show as a parity failure on the device.
// usage: ell_views n_dim waypoints n_balls n_obstacles owners      -> prints "ok ..." or "FAIL ..."; exit code 0 / 1
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "gomp_problem.h"
#include "symbolic.h"

int main(int argc, char **argv) {
    gomp_problem_config cfg;
    gomp_problem_default_config(&cfg);
    if (argc > 1) cfg.n_dim = std::atoi(argv[1]);
    if (argc > 2) cfg.waypoints = std::atoi(argv[2]);
    if (argc > 3) cfg.n_balls = std::atoi(argv[3]);
    if (argc > 4) cfg.n_obstacles = std::atoi(argv[4]);
    const int owners = argc > 5 ? std::atoi(argv[5]) : 192;
    int64_t n, m, nnzA, nnzP;
    if (gomp_problem_dims(&cfg, &n, &m, &nnzA, &nnzP)) return 2;
    std::vector<int64_t> Ap(n + 1), Ai(nnzA), Pp(n + 1), Pi(nnzP), order(n);
    std::vector<double> Px(nnzP);
    gomp_problem_pattern(&cfg, Ap.data(), Ai.data(), Pp.data(), Pi.data(), Px.data(), order.data());
    gomp::Symbolic s;
    if (!gomp::analyse(s, n, m, Pp.data(), Pi.data(), Px.data(), Ap.data(), Ai.data(), order.data(), 2 * cfg.n_dim)) {
        std::printf("FAIL analyse: %s\n", s.error.c_str());
        return 1;
    }
    gomp::OwnerEll o;
    gomp::build_owner_ell(s, owners, o);
    gomp::AssemblyEll a;
    gomp::build_assembly_ell(s, owners, a);
    std::mt19937_64 rng(7);
    std::uniform_real_distribution<double> U(-1.0, 1.0);
    const int np = s.n_pad, mu = s.m_used;
    // values the kernel would hold in shared memory, each array followed by the zero element surplus ELL entries point at
    std::vector<double> A((size_t)nnzA + 1), P((size_t)s.nnzPu + 1), xv((size_t)np + 2), wv((size_t)mu + 1), rho((size_t)mu + 1);
    for (auto *v : {&A, &P, &xv, &wv}) for (double &e : *v) e = U(rng);
    for (double &e : rho) e = 0.5 + std::fabs(U(rng));
    A[(size_t)nnzA] = P[(size_t)s.nnzPu] = xv[(size_t)np] = xv[(size_t)np + 1] = wv[(size_t)mu] = rho[(size_t)mu] = 0.0;
    for (int p = 0; p < np; ++p)
        if (s.pos_var[(size_t)p] < 0) xv[(size_t)p] = 0.0;
    double err = 0.0;
    long long checked = 0;
    // ---- rows: z = A x through the row view against the CSC arrays
    std::vector<double> z_ref((size_t)m, 0.0);
    for (int64_t j = 0; j < n; ++j)
        for (int64_t k = Ap[j]; k < Ap[j + 1]; ++k) z_ref[(size_t)Ai[k]] += A[(size_t)k] * xv[(size_t)s.var_pos[(size_t)j]];
    for (int i = 0; i < mu; ++i) {
        const int j = i / owners, oid = i % owners;
        double acc = 0.0;
        for (int q = 0; q < o.rlen[(size_t)j]; ++q) {
            const uint32_t e = o.row_ell[(size_t)(o.rbase[(size_t)j] + q) * owners + oid];
            acc += A[(e & 0xFFFFu) / 8] * xv[(e >> 16) / 8];
        }
        err = std::fmax(err, std::fabs(acc - z_ref[(size_t)s.used_rows[(size_t)i]]));
        ++checked;
    }
    // ---- columns: A' w and P x through the column views; every position owned exactly once
    std::vector<double> g_ref((size_t)np, 0.0), px_ref((size_t)np, 0.0);
    for (int64_t j = 0; j < n; ++j)
        for (int64_t k = Ap[j]; k < Ap[j + 1]; ++k) g_ref[(size_t)s.var_pos[(size_t)j]] += A[(size_t)k] * wv[(size_t)s.row_compact[(size_t)Ai[k]]];
    for (size_t k = 0; k < s.Pu_ij.size(); ++k) {
        const int pi = (int)(s.Pu_ij[k] >> 16), pj = (int)(s.Pu_ij[k] & 0xFFFFu);
        px_ref[(size_t)pi] += P[k] * xv[(size_t)pj];
        if (pi != pj) px_ref[(size_t)pj] += P[k] * xv[(size_t)pi];
    }
    std::vector<int> owned((size_t)np, 0);
    for (int c = 0; c < o.cpt; ++c)
        for (int oid = 0; oid < owners; ++oid) {
            const uint16_t p = o.col_pos[(size_t)c * owners + oid];
            if (p == 0xFFFF) continue;
            owned[p]++;
            double acc = 0.0, px = 0.0;
            for (int q = 0; q < o.clen[(size_t)c]; ++q) {
                const uint32_t e = o.col_ell[(size_t)(o.cbase[(size_t)c] + q) * owners + oid];
                acc += A[(e & 0xFFFFu) / 8] * wv[(e >> 16) / 8];
            }
            for (int q = 0; q < o.plen[(size_t)c]; ++q) {
                const uint32_t e = o.pf_ell[(size_t)(o.pbase[(size_t)c] + q) * owners + oid];
                px += P[(e & 0xFFFFu) / 8] * xv[(e >> 16) / 8];
            }
            err = std::fmax(err, std::fmax(std::fabs(acc - g_ref[p]), std::fabs(px - px_ref[p])));
            checked += 2;
        }
    for (int p = 0; p < np; ++p)
        if (owned[(size_t)p] != 1) {
            std::printf("FAIL position %d owned %d times\n", p, owned[(size_t)p]);
            return 1;
        }
    // ---- assembly lists: S = P + sigma I + A' diag(rho) A, dense reference in position order
    const double sigma = 1e-6;
    std::vector<double> S((size_t)np * np, 0.0);
    for (size_t k = 0; k < s.Pu_ij.size(); ++k) {
        const int pi = (int)(s.Pu_ij[k] >> 16), pj = (int)(s.Pu_ij[k] & 0xFFFFu);
        S[(size_t)pi * np + pj] += P[k];
        if (pi != pj) S[(size_t)pj * np + pi] += P[k];
    }
    {
        std::vector<std::vector<std::pair<int, double>>> rows((size_t)mu);
        for (int64_t j = 0; j < n; ++j)
            for (int64_t k = Ap[j]; k < Ap[j + 1]; ++k) rows[(size_t)s.row_compact[(size_t)Ai[k]]].push_back({s.var_pos[(size_t)j], A[(size_t)k]});
        for (int i = 0; i < mu; ++i)
            for (auto &e1 : rows[(size_t)i])
                for (auto &e2 : rows[(size_t)i]) S[(size_t)e1.first * np + e2.first] += rho[(size_t)i] * e1.second * e2.second;
    }
    for (int p = 0; p < np; ++p) {
        if (s.pos_var[(size_t)p] >= 0) S[(size_t)p * np + p] += sigma;
        else S[(size_t)p * np + p] = 1.0; // padding positions: identity
    }
    const int T = s.T, h = T / 2, dpb = s.dpb, bs = s.bs;
    auto contrib = [&](const std::vector<uint64_t> &con, size_t base, int maxc) {
        double acc = 0.0;
        for (int k = 0; k < maxc; ++k) {
            const uint64_t w = con[base + (size_t)k * owners];
            acc += rho[(w & 0xFFFFu) / 8] * A[((w >> 16) & 0xFFFFu) / 8] * A[((w >> 32) & 0xFFFFu) / 8];
        }
        return acc;
    };
    std::vector<int> seen((size_t)T * dpb, 0);
    for (int R = 0; R < a.rounds; ++R)
        for (int w = 0; w < a.ept; ++w)
            for (int oid = 0; oid < owners; ++oid) {
                const uint32_t meta = a.meta[(size_t)(R * a.ept + w) * owners + oid];
                if (!((meta >> 18) & 1u)) continue;
                const int e = oid + w * owners, which = e >= dpb ? 1 : 0, le = e - which * dpb;
                const int t = R == h ? h : (which == 0 ? R : T - 1 - R);
                double acc = contrib(a.con, ((size_t)(R * a.ept + w) * a.maxc) * owners + oid, a.maxc) + P[(meta & 0xFFFFu) / 8];
                const uint32_t kind = (meta >> 16) & 3u;
                if (kind == 1u) acc += sigma;
                else if (kind == 2u) acc = 1.0;
                int r = 0;
                while ((r + 1) * (r + 2) / 2 <= le) ++r;
                const int c = le - r * (r + 1) / 2;
                err = std::fmax(err, std::fabs(acc - S[(size_t)(t * bs + r) * np + (t * bs + c)]));
                seen[(size_t)t * dpb + le]++;
                ++checked;
            }
    for (size_t e = 0; e < seen.size(); ++e)
        if (seen[e] != 1) {
            std::printf("FAIL diagonal-block entry %zu assembled %d times\n", e, seen[e]);
            return 1;
        }
    // off-diagonal blocks: ELL entry e = (t, row r of stage t+1, slot q) couples position (t+1) bs + r with t bs + off_col
    const int noff = std::max(0, T - 1) * bs * s.ow;
    for (int e = 0; e < noff; ++e) {
        const int w = e / owners, oid = e % owners;
        const uint32_t meta = a.ometa[(size_t)w * owners + oid];
        const double acc = contrib(a.ocon, ((size_t)w * a.maxco) * owners + oid, a.maxco) + P[(meta & 0xFFFFu) / 8];
        const int t = e / (bs * s.ow), r = (e / s.ow) % bs;
        const int c = s.off_col[(size_t)e];
        const double want = c >= 0 ? S[(size_t)((t + 1) * bs + r) * np + (t * bs + c)] : 0.0;
        err = std::fmax(err, std::fabs(acc - want));
        ++checked;
    }
    // nothing of S lies outside the block tridiagonal the lists cover
    double outside = 0.0;
    for (int p = 0; p < np; ++p)
        for (int q = 0; q < np; ++q)
            if (std::abs(p / bs - q / bs) > 1) outside = std::fmax(outside, std::fabs(S[(size_t)p * np + q]));
    const bool ok = err < 1e-12 && outside == 0.0;
    std::printf("%s N=%d W=%d owners=%d rpt=%d cpt=%d maxc=%d maxco=%d checked=%lld max_err=%.3e outside=%.1e\n", ok ? "ok" : "FAIL", cfg.n_dim,
                cfg.waypoints, owners, o.rpt, o.cpt, a.maxc, a.maxco, checked, err, outside);
    return ok ? 0 : 1;
}
