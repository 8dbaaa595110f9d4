// emulate.cpp -- sequential CPU walk-through of the device algorithm (TEST INFRASTRUCTURE).
//
// Uses the product's symbolic analysis (osqp-solver_b200/csrc/symbolic.h) and performs, one
// phase after another and without any parallelism, exactly the steps the CUDA kernel in
// osqp-solver_b200/csrc/gomp_cuda.cu performs per QP: row compaction, Ruiz scaling on the
// CSC-ordered value array, rho classification, destination-driven assembly of the
// block-tridiagonal reduced KKT matrix, two-way block factorisation with explicit pivot-block
// inverses, the reduced-form ADMM iteration, OSQP's termination / infeasibility / adaptive-rho
// logic.  It exists so that the index maps and the algorithm can be validated against the
// oracle on a machine without a GPU; it is never linked into the product libraries.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "symbolic.h"

namespace {

constexpr double RHO_MIN = 1e-6, RHO_MAX = 1e6, RHO_EQ = 1e3, RHO_TOL = 1e-4;
constexpr double MIN_SCALING = 1e-4, MAX_SCALING = 1e4, INFTY = 1e30, DIV_TOL = 1.0 / 1e30;

struct Settings {
    double rho, sigma, alpha;
    int scaling, adaptive_rho, adaptive_rho_interval;
    double adaptive_rho_tolerance;
    int max_iter;
    double eps_abs, eps_rel, eps_prim_inf, eps_dual_inf;
    int scaled_termination, check_termination, warm_start;
};

double limit(double v) {
    v = v < MIN_SCALING ? 1.0 : v;
    return v > MAX_SCALING ? MAX_SCALING : v;
}
int classify(double l, double u) {
    if (l < -INFTY * MIN_SCALING && u > INFTY * MIN_SCALING) return -1;
    if (u - l < RHO_TOL) return 1;
    return 0;
}

struct QP {
    const gomp::Symbolic *S;
    Settings st;
    std::vector<double> A_raw, l_raw, u_raw, q_raw, x0;
    bool has_q = false, ws_pending = false, cold = true;
    // persistent (scaled) iterates
    std::vector<double> x, z, y;
    double rho = 0.1;
    int rho_updates = 0;
    // outputs
    std::vector<double> x_out, y_out;
    int status = -10, iters = 0, nfactor = 0;
    double pri_res = 0, dua_res = 0;

    // per-solve working set ("shared memory" / "registers" of the kernel)
    std::vector<double> sA, sP, sW, sRHS, Dinv, Off, l, u, D, E, qs;
    std::vector<int> ctype;
    double c = 1, cinv = 1;

    double rho_of(int i) const { return ctype[i] == -1 ? RHO_MIN : (ctype[i] == 1 ? RHO_EQ * rho : rho); }
    double rinv_of(int i) const { return 1.0 / rho_of(i); }

    template <typename F>
    void forRow(int i, F f) const {
        for (int j = 0; j < S->row_len[i]; ++j) {
            uint32_t e = S->row_ent[(size_t)j * S->m_used + i];
            f((int)(e >> 16), (int)(e & 0xFFFF));
        }
    }
    template <typename F>
    void forCol(int p, F f) const {
        for (int j = 0; j < S->col_len[p]; ++j) f((int)S->col_row[(size_t)j * S->n_pad + p], (int)(S->col_start[p] + j));
    }
    template <typename F>
    void forPfull(int p, F f) const {
        for (int j = 0; j < S->pf_len[p]; ++j) {
            uint32_t e = S->pf_ent[(size_t)j * S->n_pad + p];
            f((int)(e >> 16), (int)(e & 0xFFFF));
        }
    }

    void scale() {
        const int np = S->n_pad, mu = S->m_used;
        D.assign(np, 1.0);
        E.assign(mu, 1.0);
        c = 1.0;
        // As the device kernels do it: rn holds the rows' max norms as they stand; a pass takes the column norms, then scales
        // every entry in one row-view traversal -- (a * e_row) * d_col, OSQP's row scaling followed by its column scaling --
        // and takes the next pass's row norms from the values it writes (two traversals of A per pass instead of four).
        std::vector<double> dt(np), et(mu), rn(mu, 0.0);
        if (st.scaling > 0)
            for (int i = 0; i < mu; ++i) forRow(i, [&](int, int k) { rn[i] = std::max(rn[i], std::fabs(sA[k])); });
        for (int it = 0; it < st.scaling; ++it) {
            for (int p = 0; p < np; ++p) {
                double nrm = 0;
                forPfull(p, [&](int, int k) { nrm = std::max(nrm, std::fabs(sP[k])); });
                forCol(p, [&](int, int k) { nrm = std::max(nrm, std::fabs(sA[k])); });
                dt[p] = 1.0 / std::sqrt(limit(nrm));
            }
            for (int i = 0; i < mu; ++i) {
                et[i] = 1.0 / std::sqrt(limit(rn[i]));
                double nrm = 0;
                forRow(i, [&](int pos, int k) {
                    sA[k] = (sA[k] * et[i]) * dt[pos];
                    nrm = std::max(nrm, std::fabs(sA[k]));
                });
                rn[i] = nrm;
            }
            for (size_t k = 0; k < sP.size(); ++k) {
                uint32_t ij = S->Pu_ij[k];
                sP[k] = (sP[k] * dt[ij >> 16]) * dt[ij & 0xFFFF];
            }
            for (int p = 0; p < np; ++p) {
                qs[p] *= dt[p];
                D[p] *= dt[p];
            }
            for (int i = 0; i < mu; ++i) E[i] *= et[i];
            double sum = 0, qn = 0;
            for (int p = 0; p < np; ++p) {
                double nrm = 0;
                forPfull(p, [&](int, int k) { nrm = std::max(nrm, std::fabs(sP[k])); });
                sum += nrm;
                qn = std::max(qn, std::fabs(qs[p]));
            }
            double ct = std::max(sum / (double)S->n, limit(qn));
            ct = 1.0 / limit(ct);
            for (auto &v : sP) v *= ct;
            for (auto &v : qs) v *= ct;
            c *= ct;
        }
        cinv = 1.0 / c;
        for (int i = 0; i < mu; ++i) {
            l[i] *= E[i];
            u[i] *= E[i];
        }
    }

    // in-place inverse of the SPD block stored packed-lower at M (dpb entries)
    void invertPacked(double *M) {
        const int bs = S->bs;
        std::vector<double> F((size_t)bs * bs);
        for (int r = 0; r < bs; ++r)
            for (int cc = 0; cc <= r; ++cc) F[r * bs + cc] = F[cc * bs + r] = M[r * (r + 1) / 2 + cc];
        for (int k = 0; k < bs; ++k) { // Gauss-Jordan, no pivoting (SPD)
            double pinv = 1.0 / F[k * bs + k];
            for (int i = 0; i < bs; ++i) {
                if (i == k) continue;
                double f = F[i * bs + k] * pinv;
                for (int j = 0; j < bs; ++j)
                    if (j != k) F[i * bs + j] -= f * F[k * bs + j];
                F[i * bs + k] = -f;
            }
            for (int j = 0; j < bs; ++j)
                if (j != k) F[k * bs + j] *= pinv;
            F[k * bs + k] = pinv;
        }
        for (int r = 0; r < bs; ++r)
            for (int cc = 0; cc <= r; ++cc) M[r * (r + 1) / 2 + cc] = F[r * bs + cc];
    }
    double dinv(int t, int r, int cc) const {
        return r >= cc ? Dinv[(size_t)t * S->dpb + r * (r + 1) / 2 + cc] : Dinv[(size_t)t * S->dpb + cc * (cc + 1) / 2 + r];
    }
    // D_t -= Off_{t-1} Dinv_{t-1} Off_{t-1}'   (contribution of the block above)
    void schurFromAbove(int t) {
        const int bs = S->bs, ow = S->ow;
        for (int r = 0; r < bs; ++r)
            for (int cc = 0; cc <= r; ++cc) {
                double acc = 0;
                for (int s = 0; s < ow; ++s) {
                    int a = S->off_col[((size_t)(t - 1) * bs + r) * ow + s];
                    if (a < 0) continue;
                    for (int s2 = 0; s2 < ow; ++s2) {
                        int b = S->off_col[((size_t)(t - 1) * bs + cc) * ow + s2];
                        if (b < 0) continue;
                        acc += Off[((size_t)(t - 1) * bs + r) * ow + s] * dinv(t - 1, a, b) * Off[((size_t)(t - 1) * bs + cc) * ow + s2];
                    }
                }
                Dinv[(size_t)t * S->dpb + r * (r + 1) / 2 + cc] -= acc;
            }
    }
    // D_t -= Off_t' Dinv_{t+1} Off_t   (contribution of the block below)
    void schurFromBelow(int t) {
        const int bs = S->bs, owT = S->owT;
        for (int r = 0; r < bs; ++r)
            for (int cc = 0; cc <= r; ++cc) {
                double acc = 0;
                for (int s = 0; s < owT; ++s) {
                    int a = S->offT_row[((size_t)t * bs + r) * owT + s];
                    if (a < 0) continue;
                    double va = Off[(size_t)t * bs * S->ow + S->offT_slot[((size_t)t * bs + r) * owT + s]];
                    for (int s2 = 0; s2 < owT; ++s2) {
                        int b = S->offT_row[((size_t)t * bs + cc) * owT + s2];
                        if (b < 0) continue;
                        double vb = Off[(size_t)t * bs * S->ow + S->offT_slot[((size_t)t * bs + cc) * owT + s2]];
                        acc += va * dinv(t + 1, a, b) * vb;
                    }
                }
                Dinv[(size_t)t * S->dpb + r * (r + 1) / 2 + cc] -= acc;
            }
    }
    void factor() {
        const int bs = S->bs, T = S->T, ow = S->ow, dpb = S->dpb;
        std::vector<double> rv(S->m_used);
        for (int i = 0; i < S->m_used; ++i) rv[i] = rho_of(i);
        for (size_t e = 0; e < (size_t)T * dpb; ++e) {
            double acc = 0;
            for (uint32_t q = S->sd_ptr[e]; q < S->sd_ptr[e + 1]; ++q) {
                const gomp::Contrib &cn = S->sd_con[q];
                acc += rv[cn.row] * sA[cn.k1] * sA[cn.k2];
            }
            if (S->sd_pk[e] != 0xFFFF) acc += sP[S->sd_pk[e]];
            int t = (int)(e / dpb), loc = (int)(e % dpb);
            int r = (int)((std::sqrt(8.0 * loc + 1) - 1) / 2);
            while (r * (r + 1) / 2 > loc) --r;
            while ((r + 1) * (r + 2) / 2 <= loc) ++r;
            int cc = loc - r * (r + 1) / 2;
            if (r == cc) acc = S->pos_var[(size_t)t * bs + r] < 0 ? 1.0 : acc + st.sigma;
            Dinv[e] = acc;
        }
        for (size_t e = 0; e < (size_t)std::max(0, T - 1) * bs * ow; ++e) {
            double acc = 0;
            for (uint32_t q = S->so_ptr[e]; q < S->so_ptr[e + 1]; ++q) {
                const gomp::Contrib &cn = S->so_con[q];
                acc += rv[cn.row] * sA[cn.k1] * sA[cn.k2];
            }
            if (S->so_pk[e] != 0xFFFF) acc += sP[S->so_pk[e]];
            Off[e] = acc;
        }
        const int h = T / 2;
        for (int t = 0; t < h; ++t) {
            if (t > 0) schurFromAbove(t);
            invertPacked(&Dinv[(size_t)t * dpb]);
        }
        for (int t = T - 1; t > h; --t) {
            if (t < T - 1) schurFromBelow(t);
            invertPacked(&Dinv[(size_t)t * dpb]);
        }
        if (h > 0) schurFromAbove(h);
        if (h < T - 1) schurFromBelow(h);
        invertPacked(&Dinv[(size_t)h * dpb]);
        nfactor++;
    }
    // helpers on block vectors held in sRHS
    void offMul(int t, const double *vin /*block t*/, double *acc /*block t+1, subtract*/) const {
        const int bs = S->bs, ow = S->ow;
        for (int r = 0; r < bs; ++r)
            for (int s = 0; s < ow; ++s) {
                int a = S->off_col[((size_t)t * bs + r) * ow + s];
                if (a >= 0) acc[r] -= Off[((size_t)t * bs + r) * ow + s] * vin[a];
            }
    }
    void offTMul(int t, const double *vin /*block t+1*/, double *acc /*block t, subtract*/) const {
        const int bs = S->bs, owT = S->owT;
        for (int cc = 0; cc < bs; ++cc)
            for (int s = 0; s < owT; ++s) {
                int r = S->offT_row[((size_t)t * bs + cc) * owT + s];
                if (r >= 0) acc[cc] -= Off[(size_t)t * bs * S->ow + S->offT_slot[((size_t)t * bs + cc) * owT + s]] * vin[r];
            }
    }
    void gemv(int t, const double *vin, double *vout) const {
        const int bs = S->bs;
        for (int r = 0; r < bs; ++r) {
            double acc = 0;
            for (int cc = 0; cc < bs; ++cc) acc += dinv(t, r, cc) * vin[cc];
            vout[r] = acc;
        }
    }
    void blockSolve() { // in place on sRHS
        const int bs = S->bs, T = S->T, h = T / 2;
        std::vector<double> yv(bs), tmp(bs);
        double *b = sRHS.data();
        for (int t = 0; t < h; ++t) { // top chain: b_t <- u_t = Dinv_t y_t
            for (int r = 0; r < bs; ++r) yv[r] = b[t * bs + r];
            if (t > 0) offMul(t - 1, &b[(t - 1) * bs], yv.data());
            gemv(t, yv.data(), &b[t * bs]);
        }
        for (int t = T - 1; t > h; --t) { // bottom chain
            for (int r = 0; r < bs; ++r) yv[r] = b[t * bs + r];
            if (t < T - 1) offTMul(t, &b[(t + 1) * bs], yv.data());
            gemv(t, yv.data(), &b[t * bs]);
        }
        for (int r = 0; r < bs; ++r) yv[r] = b[h * bs + r];
        if (h > 0) offMul(h - 1, &b[(h - 1) * bs], yv.data());
        if (h < T - 1) offTMul(h, &b[(h + 1) * bs], yv.data());
        gemv(h, yv.data(), &b[h * bs]);
        for (int t = h - 1; t >= 0; --t) { // x_t = u_t - Dinv_t Off_t' x_{t+1}
            for (int r = 0; r < bs; ++r) yv[r] = 0;
            offTMul(t, &b[(t + 1) * bs], yv.data());
            gemv(t, yv.data(), tmp.data());
            for (int r = 0; r < bs; ++r) b[t * bs + r] += tmp[r];
        }
        for (int t = h + 1; t < T; ++t) { // x_t = u_t - Dinv_t Off_{t-1} x_{t-1}
            for (int r = 0; r < bs; ++r) yv[r] = 0;
            offMul(t - 1, &b[(t - 1) * bs], yv.data());
            gemv(t, yv.data(), tmp.data());
            for (int r = 0; r < bs; ++r) b[t * bs + r] += tmp[r];
        }
    }

    void solve() {
        const gomp::Symbolic &Y = *S;
        const int np = Y.n_pad, mu = Y.m_used;
        const double NaN = std::numeric_limits<double>::quiet_NaN();
        // ---- load
        sA = A_raw;
        sP = Y.Pu_val;
        l.resize(mu);
        u.resize(mu);
        for (int i = 0; i < mu; ++i) {
            l[i] = std::max(l_raw[Y.used_rows[i]], -INFTY);
            u[i] = std::min(u_raw[Y.used_rows[i]], INFTY);
        }
        bool empty_infeasible = false;
        for (int e : Y.empty_rows)
            if (!(l_raw[e] <= 0.0 && 0.0 <= u_raw[e])) empty_infeasible = true;
        qs.assign(np, 0.0);
        if (has_q)
            for (int p = 0; p < np; ++p)
                if (Y.pos_var[p] >= 0) qs[p] = q_raw[Y.pos_var[p]];
        sW.assign(mu, 0.0);
        sRHS.assign(np, 0.0);
        Dinv.assign((size_t)Y.T * Y.dpb, 0.0);
        Off.assign((size_t)std::max(1, (Y.T - 1) * Y.bs * Y.ow), 0.0);
        D.assign(np, 1.0);
        E.assign(mu, 1.0);
        c = cinv = 1.0;
        if (st.scaling > 0) scale();
        ctype.resize(mu);
        for (int i = 0; i < mu; ++i) ctype[i] = classify(l[i], u[i]);
        // ---- iterates
        if (cold) {
            x.assign(np, 0.0);
            z.assign(mu, 0.0);
            y.assign(mu, 0.0);
            rho = std::min(std::max(st.rho, RHO_MIN), RHO_MAX);
            rho_updates = 0;
            cold = false;
        }
        if (ws_pending) {
            for (int p = 0; p < np; ++p) x[p] = Y.pos_var[p] >= 0 ? (1.0 / D[p]) * x0[Y.pos_var[p]] : 0.0;
            for (int i = 0; i < mu; ++i) {
                double acc = 0;
                forRow(i, [&](int pos, int k) { acc += sA[k] * x[pos]; });
                z[i] = acc;
            }
            ws_pending = false;
        }
        status = -10;
        if (empty_infeasible) { // an empty row whose bounds exclude 0: trivially infeasible
            status = -3;
            iters = 0;
        }
        factor();
        const double alpha = st.alpha, sigma = st.sigma;
        std::vector<double> dx(np, 0.0), dy(mu, 0.0), Ax(mu), Px(np), Aty(np), pri(mu), dua(np);
        int iter = 0;
        bool can_check = false;
        auto update_info = [&]() {
            for (int i = 0; i < mu; ++i) {
                double acc = 0;
                forRow(i, [&](int pos, int k) { acc += sA[k] * x[pos]; });
                Ax[i] = acc;
                pri[i] = Ax[i] - z[i];
            }
            for (int p = 0; p < np; ++p) {
                double px = 0, aty = 0;
                forPfull(p, [&](int other, int k) { px += sP[k] * x[other]; });
                forCol(p, [&](int row, int k) { aty += sA[k] * y[row]; });
                Px[p] = px;
                Aty[p] = aty;
                dua[p] = qs[p] + px + aty;
            }
            pri_res = 0;
            dua_res = 0;
            const bool uns = st.scaling > 0 && !st.scaled_termination;
            for (int i = 0; i < mu; ++i) pri_res = std::max(pri_res, std::fabs(uns ? (1.0 / E[i]) * pri[i] : pri[i]));
            for (int p = 0; p < np; ++p) dua_res = std::max(dua_res, std::fabs(uns ? (1.0 / D[p]) * dua[p] : dua[p]));
            if (uns) dua_res *= cinv;
        };
        auto norm_scaled = [&](const std::vector<double> &v, const std::vector<double> &sc, bool inv, bool uns) {
            double r = 0;
            for (size_t i = 0; i < v.size(); ++i) r = std::max(r, std::fabs(uns ? (inv ? 1.0 / sc[i] : sc[i]) * v[i] : v[i]));
            return r;
        };
        auto primal_infeasible = [&](double eps) {
            const bool uns = st.scaling > 0 && !st.scaled_termination;
            for (int i = 0; i < mu; ++i) {
                if (u[i] > INFTY * MIN_SCALING) {
                    if (l[i] < -INFTY * MIN_SCALING) dy[i] = 0.0;
                    else dy[i] = std::min(dy[i], 0.0);
                } else if (l[i] < -INFTY * MIN_SCALING)
                    dy[i] = std::max(dy[i], 0.0);
            }
            double nrm = norm_scaled(dy, E, false, uns);
            if (nrm > DIV_TOL) {
                double lhs = 0;
                for (int i = 0; i < mu; ++i) lhs += u[i] * std::max(dy[i], 0.0) + l[i] * std::min(dy[i], 0.0);
                if (lhs < -eps * nrm) {
                    double r = 0;
                    for (int p = 0; p < np; ++p) {
                        double acc = 0;
                        forCol(p, [&](int row, int k) { acc += sA[k] * dy[row]; });
                        r = std::max(r, std::fabs(uns ? (1.0 / D[p]) * acc : acc));
                    }
                    return r < eps * nrm;
                }
            }
            return false;
        };
        auto dual_infeasible = [&](double eps) {
            const bool uns = st.scaling > 0 && !st.scaled_termination;
            double nrm = norm_scaled(dx, D, false, uns), cs = uns ? c : 1.0;
            if (nrm > DIV_TOL) {
                double qdx = 0;
                for (int p = 0; p < np; ++p) qdx += qs[p] * dx[p];
                if (qdx < -cs * eps * nrm) {
                    double r = 0;
                    for (int p = 0; p < np; ++p) {
                        double acc = 0;
                        forPfull(p, [&](int other, int k) { acc += sP[k] * dx[other]; });
                        r = std::max(r, std::fabs(uns ? (1.0 / D[p]) * acc : acc));
                    }
                    if (r < cs * eps * nrm) {
                        for (int i = 0; i < mu; ++i) {
                            double acc = 0;
                            forRow(i, [&](int pos, int k) { acc += sA[k] * dx[pos]; });
                            if (uns) acc *= 1.0 / E[i];
                            if ((u[i] < INFTY * MIN_SCALING && acc > eps * nrm) || (l[i] > -INFTY * MIN_SCALING && acc < -eps * nrm))
                                return false;
                        }
                        return true;
                    }
                }
            }
            return false;
        };
        auto check_termination = [&](bool approx) {
            double ea = st.eps_abs, er = st.eps_rel, epi = st.eps_prim_inf, edi = st.eps_dual_inf;
            if (pri_res > INFTY || dua_res > INFTY) {
                status = -7;
                return true;
            }
            if (approx) {
                ea *= 10; er *= 10; epi *= 10; edi *= 10;
            }
            const bool uns = st.scaling > 0 && !st.scaled_termination;
            bool prc = false, drc = false, pic = false, dic = false;
            if (mu == 0)
                prc = true;
            else {
                double eps_prim = ea + er * std::max(norm_scaled(z, E, true, uns), norm_scaled(Ax, E, true, uns));
                if (pri_res < eps_prim) prc = true;
                else pic = primal_infeasible(epi);
            }
            double mx = std::max(norm_scaled(qs, D, true, uns), std::max(norm_scaled(Aty, D, true, uns), norm_scaled(Px, D, true, uns)));
            if (uns) mx *= cinv;
            double eps_dual = ea + er * mx;
            if (dua_res < eps_dual) drc = true;
            else dic = dual_infeasible(edi);
            if (prc && drc) {
                status = approx ? 2 : 1;
                return true;
            }
            if (pic) {
                status = approx ? 3 : -3;
                return true;
            }
            if (dic) {
                status = approx ? 4 : -4;
                return true;
            }
            return false;
        };
        auto rho_estimate = [&]() {
            double p = norm_scaled(pri, E, false, false), d = norm_scaled(dua, D, false, false);
            double pn = std::max(norm_scaled(z, E, false, false), norm_scaled(Ax, E, false, false));
            double dn = std::max(norm_scaled(qs, D, false, false), std::max(norm_scaled(Aty, D, false, false), norm_scaled(Px, D, false, false)));
            p /= (pn + DIV_TOL);
            d /= (dn + DIV_TOL);
            double est = rho * std::sqrt(p / d);
            return std::min(std::max(est, RHO_MIN), RHO_MAX);
        };

        if (status == -10) {
            for (iter = 1; iter <= st.max_iter; ++iter) {
                for (int i = 0; i < mu; ++i) sW[i] = rho_of(i) * z[i] - y[i];
                for (int p = 0; p < np; ++p) {
                    double acc = sigma * x[p] - qs[p];
                    forCol(p, [&](int row, int k) { acc += sA[k] * sW[row]; });
                    sRHS[p] = acc;
                }
                blockSolve();
                for (int p = 0; p < np; ++p) {
                    double xn = alpha * sRHS[p] + (1.0 - alpha) * x[p];
                    dx[p] = xn - x[p];
                    x[p] = xn;
                }
                for (int i = 0; i < mu; ++i) {
                    double zt = 0;
                    forRow(i, [&](int pos, int k) { zt += sA[k] * sRHS[pos]; });
                    double zr = alpha * zt + (1.0 - alpha) * z[i];
                    double zn = std::min(std::max(zr + rinv_of(i) * y[i], l[i]), u[i]);
                    double d = rho_of(i) * (zr - zn);
                    dy[i] = d;
                    y[i] += d;
                    z[i] = zn;
                }
                can_check = st.check_termination && (iter % st.check_termination == 0);
                if (can_check) {
                    update_info();
                    iters = iter;
                    if (check_termination(false)) break;
                }
                if (st.adaptive_rho && st.adaptive_rho_interval && iter % st.adaptive_rho_interval == 0) {
                    if (!can_check) {
                        update_info();
                        iters = iter;
                    }
                    double rn = rho_estimate();
                    if (rn > rho * st.adaptive_rho_tolerance || rn < rho / st.adaptive_rho_tolerance) {
                        rho = rn;
                        rho_updates++;
                        factor();
                    }
                }
            }
            if (!can_check) {
                update_info();
                iters = iter - 1;
                check_termination(false);
            }
            if (status == -10 && !check_termination(true)) status = -2;
        }
        // ---- store
        x_out.assign(Y.n, 0.0);
        y_out.assign(Y.m, 0.0);
        const bool infeasible = status == -3 || status == 3 || status == -4 || status == 4 || status == -7;
        if (!infeasible) {
            for (int p = 0; p < np; ++p)
                if (Y.pos_var[p] >= 0) x_out[Y.pos_var[p]] = D[p] * x[p];
            for (int i = 0; i < mu; ++i) y_out[Y.used_rows[i]] = (E[i] * y[i]) * cinv;
        } else {
            std::fill(x_out.begin(), x_out.end(), NaN);
            std::fill(y_out.begin(), y_out.end(), NaN);
            x.assign(np, 0.0);
            z.assign(mu, 0.0);
            y.assign(mu, 0.0);
        }
    }
};

struct Batch {
    gomp::Symbolic S;
    std::vector<QP> qps;
};

int exit_code(int status) {
    switch (status) {
    case 1: return 0;
    case 2: return 3;
    case -3: return 1;
    case 3: return 4;
    case -4: return 2;
    case 4: return 5;
    case -2: return 6;
    case -7: return 9;
    default: return 10;
    }
}

} // namespace

extern "C" {

void *emul_create(int64_t batch, int64_t n, int64_t m, const int64_t *Pp, const int64_t *Pi, const double *Px,
                  const int64_t *Ap, const int64_t *Ai, const double *settings15, const int64_t *var_order, int block_size,
                  char *err, int errlen) {
    Batch *B = new Batch();
    if (!gomp::analyse(B->S, n, m, Pp, Pi, Px, Ap, Ai, var_order, block_size)) {
        if (err) std::strncpy(err, B->S.error.c_str(), (size_t)errlen - 1);
        delete B;
        return nullptr;
    }
    Settings st;
    st.rho = settings15[0]; st.sigma = settings15[1]; st.alpha = settings15[2];
    st.scaling = (int)settings15[3]; st.adaptive_rho = (int)settings15[4]; st.adaptive_rho_interval = (int)settings15[5];
    st.adaptive_rho_tolerance = settings15[6]; st.max_iter = (int)settings15[7];
    st.eps_abs = settings15[8]; st.eps_rel = settings15[9]; st.eps_prim_inf = settings15[10]; st.eps_dual_inf = settings15[11];
    st.scaled_termination = (int)settings15[12]; st.check_termination = (int)settings15[13]; st.warm_start = (int)settings15[14];
    B->qps.resize((size_t)batch);
    for (auto &q : B->qps) {
        q.S = &B->S;
        q.st = st;
    }
    return B;
}
void emul_destroy(void *h) { delete (Batch *)h; }

// cold = 1: gomp_batch_setup semantics; cold = 0: gomp_batch_update semantics
int emul_set_data(void *h, const double *Ax, const double *l, const double *u, const double *q, int cold) {
    Batch *B = (Batch *)h;
    const gomp::Symbolic &S = B->S;
    for (size_t b = 0; b < B->qps.size(); ++b) {
        QP &Q = B->qps[b];
        if (Ax) Q.A_raw.assign(Ax + b * S.nnzA, Ax + (b + 1) * S.nnzA);
        if (l) Q.l_raw.assign(l + b * S.m, l + (b + 1) * S.m);
        if (u) Q.u_raw.assign(u + b * S.m, u + (b + 1) * S.m);
        if (q) {
            Q.q_raw.assign(q + b * S.n, q + (b + 1) * S.n);
            Q.has_q = true;
        }
        if (cold) {
            Q.cold = true;
            Q.ws_pending = false;
        }
        for (int64_t i = 0; i < S.m; ++i)
            if (Q.l_raw[(size_t)i] > Q.u_raw[(size_t)i]) return 1;
    }
    return 0;
}
void emul_warm_start_x(void *h, const double *x0) {
    Batch *B = (Batch *)h;
    for (size_t b = 0; b < B->qps.size(); ++b) {
        B->qps[b].x0.assign(x0 + b * B->S.n, x0 + (b + 1) * B->S.n);
        B->qps[b].ws_pending = true;
    }
}
void emul_solve(void *h) {
    Batch *B = (Batch *)h;
    for (auto &q : B->qps) q.solve();
}
void emul_get(void *h, double *x, double *y, int32_t *code, int32_t *iters, int32_t *rho_updates, double *rho, int32_t *nfactor) {
    Batch *B = (Batch *)h;
    const gomp::Symbolic &S = B->S;
    for (size_t b = 0; b < B->qps.size(); ++b) {
        const QP &Q = B->qps[b];
        if (x) std::memcpy(x + b * S.n, Q.x_out.data(), (size_t)S.n * sizeof(double));
        if (y) std::memcpy(y + b * S.m, Q.y_out.data(), (size_t)S.m * sizeof(double));
        if (code) code[b] = exit_code(Q.status);
        if (iters) iters[b] = Q.iters;
        if (rho_updates) rho_updates[b] = Q.rho_updates;
        if (rho) rho[b] = Q.rho;
        if (nfactor) nfactor[b] = Q.nfactor;
    }
}
void emul_dims(void *h, int32_t *out8) {
    const gomp::Symbolic &S = ((Batch *)h)->S;
    out8[0] = S.m_used; out8[1] = S.bs; out8[2] = S.T; out8[3] = S.n_pad;
    out8[4] = S.roww; out8[5] = S.colw; out8[6] = S.ow; out8[7] = S.owT;
}
}
