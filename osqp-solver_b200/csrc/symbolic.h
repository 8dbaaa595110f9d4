// symbolic.h -- host-side symbolic analysis for the batched reduced-KKT ADMM kernels.
//
// One analysis per sparsity pattern, shared by every QP of a batch (the role AMD + the QDLDL
// etree/symbolic phase play inside every OSQP setup, reference src/osqp-wrapper.h:28 ->
// osqp_setup).  It turns (pattern of P, pattern of A, variable order) into
// the index maps the device kernels stream:
//   * row compaction: rows of A without any entry never leave z = y = 0 and are dropped from
//     the on-chip problem (GOMP over-allocates ~44% of its rows);
//   * a variable order under which S = P + sigma I + A' diag(rho) A is block tridiagonal with
//     bs x bs blocks (GOMP: stage-interleaved (q_t, v_t), bs = 2N);
//   * ELL-transposed row and column views of A whose entries point into the caller's
//     CSC-ordered value array (so values are staged once, unpermuted);
//   * full-symmetric and upper-triangular views of P;
//   * "assembly lists": for every stored entry of S (packed lower diagonal blocks, ELL
//     off-diagonal blocks) the list of (row, A-entry, A-entry) products and the P entry that
//     sum into it -- destination-driven, hence deterministic and atomic-free.
// Pure C++17, no CUDA: also compiled into the CPU emulator used by the tests.
#ifndef GOMP_SYMBOLIC_H
#define GOMP_SYMBOLIC_H

#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstdlib>
#include <queue>
#include <string>
#include <vector>

namespace gomp {

struct Contrib { // rho[row] * A[k1] * A[k2]
    uint16_t row, k1, k2;
};

struct Symbolic {
    int64_t n = 0, m = 0, nnzA = 0, nnzPu = 0;
    int m_used = 0, bs = 0, T = 0, n_pad = 0;
    int roww = 0, colw = 0; // ELL widths of the row / column views of A
    int ow = 0, owT = 0;    // ELL widths of off-diagonal blocks (by row of block t+1 / by column = row of block t)
    int dpb = 0;            // packed lower-triangle entries per diagonal block
    int pfw = 0;            // ELL width of the full-symmetric view of P

    std::vector<int32_t> var_pos;   // [n]      variable -> position
    std::vector<int32_t> pos_var;   // [n_pad]  position -> variable, -1 = padding
    std::vector<int32_t> used_rows; // [m_used] compact row -> original row
    std::vector<int32_t> row_compact; // [m]    original row -> compact row, -1 = empty
    std::vector<int32_t> empty_rows;  // original indices of rows without entries

    // row view of A, ELL-transposed: ent[j * m_used + i] = (pos << 16 | k), 0xFFFFFFFF = none
    std::vector<uint32_t> row_ent;
    std::vector<uint8_t> row_len; // [m_used]
    // column view by position: values of position p are A[col_start[p] .. +col_len[p]),
    // their compact rows col_row[j * n_pad + p]
    std::vector<uint32_t> col_start; // [n_pad]
    std::vector<uint8_t> col_len;    // [n_pad]
    std::vector<uint16_t> col_row;   // [colw * n_pad]
    // compact views staged into shared memory by the kernel
    std::vector<uint16_t> row_ptr;   // [m_used + 1] CSR pointers over compact rows
    std::vector<uint32_t> row_csr;   // [nnzA] (pos << 16 | k), row-major
    std::vector<uint32_t> row_range; // [m_used] (first entry << 8 | length) into row_csr
    std::vector<uint16_t> csc_row;   // [nnzA] compact row of CSC entry k
    std::vector<uint32_t> col_range; // [n_pad] (start << 8 | len) of position p in the CSC value array
    std::vector<uint16_t> pf_ptr;    // [n_pad + 1] CSR pointers of the full-symmetric view of P
    std::vector<uint32_t> pf_csr;    // (other_pos << 16 | kPu)

    // P, upper triangle in CSC order of the input
    std::vector<double> Pu_val;   // [nnzPu] raw values
    std::vector<uint32_t> Pu_ij;  // [nnzPu] (pos_i << 16 | pos_j)
    // full symmetric view by position, ELL-transposed: (other_pos << 16 | kPu)
    std::vector<uint32_t> pf_ent; // [pfw * n_pad]
    std::vector<uint8_t> pf_len;  // [n_pad]

    // S storage: diag block t entry (r >= c) at t*dpb + r(r+1)/2 + c; off block t (rows of
    // block t+1, columns of block t) slot (t*bs + r)*ow + s with column off_col[slot] (-1 none)
    std::vector<int8_t> off_col;     // [(T-1)*bs*ow]
    // transposed view: for column c of off block t: entries offT_row[(t*bs+c)*owT + s] = r (or -1),
    // value slot offT_slot[...]
    std::vector<int8_t> offT_row;    // [(T-1)*bs*owT]
    std::vector<uint16_t> offT_slot; // [(T-1)*bs*owT]  slot within block t: r*ow + s
    // assembly lists
    std::vector<uint32_t> sd_ptr;    // [T*dpb + 1]
    std::vector<Contrib> sd_con;
    std::vector<uint16_t> sd_pk;     // [T*dpb] kPu or 0xFFFF
    std::vector<uint32_t> so_ptr;    // [(T-1)*bs*ow + 1]
    std::vector<Contrib> so_con;
    std::vector<uint16_t> so_pk;     // [(T-1)*bs*ow]

    std::string error;
};

namespace detail {

// Reverse Cuthill-McKee on the adjacency lists of S; returns order[p] = variable.
inline std::vector<int64_t> rcm(int64_t n, const std::vector<std::vector<int32_t>> &adj) {
    std::vector<int64_t> order;
    order.reserve((size_t)n);
    std::vector<char> seen((size_t)n, 0);
    auto bfs_far = [&](int32_t s) { // last vertex of a BFS from s (pseudo-peripheral search)
        std::vector<int32_t> dist((size_t)n, -1);
        std::queue<int32_t> q;
        q.push(s);
        dist[s] = 0;
        int32_t last = s;
        while (!q.empty()) {
            int32_t v = q.front();
            q.pop();
            last = v;
            for (int32_t w : adj[v])
                if (dist[w] < 0 && !seen[w]) {
                    dist[w] = dist[v] + 1;
                    q.push(w);
                }
        }
        return last;
    };
    for (int64_t root = 0; root < n; ++root) {
        if (seen[root]) continue;
        int32_t s = bfs_far(bfs_far((int32_t)root));
        std::queue<int32_t> q;
        q.push(s);
        seen[s] = 1;
        while (!q.empty()) {
            int32_t v = q.front();
            q.pop();
            order.push_back(v);
            std::vector<int32_t> nb;
            for (int32_t w : adj[v])
                if (!seen[w]) {
                    seen[w] = 1;
                    nb.push_back(w);
                }
            std::sort(nb.begin(), nb.end(), [&](int32_t a, int32_t b) {
                return adj[a].size() != adj[b].size() ? adj[a].size() < adj[b].size() : a < b;
            });
            for (int32_t w : nb) q.push(w);
        }
    }
    std::reverse(order.begin(), order.end());
    return order;
}

} // namespace detail

// Returns false and fills sym.error when the structure cannot be handled.
inline bool analyse(Symbolic &sym, int64_t n, int64_t m, const int64_t *Pp, const int64_t *Pi, const double *Px,
                    const int64_t *Ap, const int64_t *Ai, const int64_t *var_order, int block_size,
                    int max_block = 32) {
    sym = Symbolic();
    sym.n = n;
    sym.m = m;
    sym.nnzA = Ap[n];
    if (n <= 0 || m < 0) return sym.error = "n must be positive and m non-negative", false;
    if (sym.nnzA > 65535) return sym.error = "nnz(A) > 65535: pattern indices are packed in 16 bits", false;

    // ---- rows -------------------------------------------------------------------------------
    std::vector<int> row_cnt((size_t)m, 0);
    for (int64_t k = 0; k < sym.nnzA; ++k) {
        if (Ai[k] < 0 || Ai[k] >= m) return sym.error = "row index of A out of range", false;
        row_cnt[(size_t)Ai[k]]++;
    }
    sym.row_compact.assign((size_t)m, -1);
    for (int64_t i = 0; i < m; ++i) {
        if (row_cnt[(size_t)i] > 0) {
            sym.row_compact[(size_t)i] = (int32_t)sym.used_rows.size();
            sym.used_rows.push_back((int32_t)i);
        } else {
            sym.empty_rows.push_back((int32_t)i);
        }
    }
    sym.m_used = (int)sym.used_rows.size();
    if (sym.m_used > 65535) return sym.error = "more than 65535 non-empty rows", false;

    // ---- upper triangle of P ------------------------------------------------------------------
    struct PE { int64_t i, j; double v; };
    std::vector<PE> pu;
    for (int64_t j = 0; j < n; ++j)
        for (int64_t k = Pp[j]; k < Pp[j + 1]; ++k) {
            if (Pi[k] < 0 || Pi[k] >= n) return sym.error = "row index of P out of range", false;
            if (Pi[k] <= j) pu.push_back(PE{Pi[k], j, Px[k]});
        }
    sym.nnzPu = (int64_t)pu.size();
    if (sym.nnzPu > 65534) return sym.error = "nnz(triu P) > 65534", false;

    // ---- adjacency of S = P + A'A (variables) --------------------------------------------------
    std::vector<std::vector<int32_t>> rows_of((size_t)m); // variables per row
    for (int64_t j = 0; j < n; ++j)
        for (int64_t k = Ap[j]; k < Ap[j + 1]; ++k) rows_of[(size_t)Ai[k]].push_back((int32_t)j);

    // ---- variable order ------------------------------------------------------------------------
    std::vector<int64_t> order((size_t)n);
    if (var_order) {
        std::vector<char> hit((size_t)n, 0);
        for (int64_t p = 0; p < n; ++p) {
            int64_t v = var_order[p];
            if (v < 0 || v >= n || hit[(size_t)v]) return sym.error = "var_order is not a permutation", false;
            hit[(size_t)v] = 1;
            order[(size_t)p] = v;
        }
    } else if (block_size > 0) {
        for (int64_t p = 0; p < n; ++p) order[(size_t)p] = p;
    } else {
        std::vector<std::vector<int32_t>> adj((size_t)n);
        for (const auto &e : pu)
            if (e.i != e.j) {
                adj[(size_t)e.i].push_back((int32_t)e.j);
                adj[(size_t)e.j].push_back((int32_t)e.i);
            }
        for (const auto &vs : rows_of)
            for (size_t a = 0; a < vs.size(); ++a)
                for (size_t b = a + 1; b < vs.size(); ++b) {
                    adj[(size_t)vs[a]].push_back(vs[b]);
                    adj[(size_t)vs[b]].push_back(vs[a]);
                }
        for (auto &l : adj) {
            std::sort(l.begin(), l.end());
            l.erase(std::unique(l.begin(), l.end()), l.end());
        }
        order = detail::rcm(n, adj);
    }
    sym.var_pos.assign((size_t)n, 0);
    for (int64_t p = 0; p < n; ++p) sym.var_pos[(size_t)order[(size_t)p]] = (int32_t)p;

    // ---- block size -----------------------------------------------------------------------------
    auto coupled_span = [&]() { // bandwidth of S under the order
        int64_t bw = 0;
        for (const auto &e : pu) bw = std::max<int64_t>(bw, std::abs((int64_t)sym.var_pos[(size_t)e.i] - sym.var_pos[(size_t)e.j]));
        for (const auto &vs : rows_of) {
            int32_t lo = INT32_MAX, hi = -1;
            for (int32_t v : vs) {
                lo = std::min(lo, sym.var_pos[(size_t)v]);
                hi = std::max(hi, sym.var_pos[(size_t)v]);
            }
            if (hi >= 0) bw = std::max<int64_t>(bw, hi - lo);
        }
        return bw;
    };
    int bs = block_size > 0 ? block_size : (int)std::max<int64_t>(1, coupled_span());
    if (bs > max_block) {
        sym.error = "block size " + std::to_string(bs) + " exceeds the supported maximum " + std::to_string(max_block) +
                    " (pass var_order/block_size for a block-tridiagonal ordering)";
        return false;
    }
    sym.bs = bs;
    sym.T = (int)((n + bs - 1) / bs);
    sym.n_pad = sym.T * bs;
    sym.dpb = bs * (bs + 1) / 2;
    if (sym.n_pad > 65535) return sym.error = "more than 65535 variables", false;
    sym.pos_var.assign((size_t)sym.n_pad, -1);
    for (int64_t v = 0; v < n; ++v) sym.pos_var[(size_t)sym.var_pos[(size_t)v]] = (int32_t)v;

    // ---- row / column views of A -----------------------------------------------------------------
    {
        std::vector<std::vector<uint32_t>> ent((size_t)sym.m_used);
        for (int64_t j = 0; j < n; ++j)
            for (int64_t k = Ap[j]; k < Ap[j + 1]; ++k)
                ent[(size_t)sym.row_compact[(size_t)Ai[k]]].push_back(((uint32_t)sym.var_pos[(size_t)j] << 16) | (uint32_t)k);
        for (auto &e : ent) {
            std::sort(e.begin(), e.end());
            sym.roww = std::max(sym.roww, (int)e.size());
        }
        if (sym.roww > 255) return sym.error = "a row of A has more than 255 entries", false;
        sym.row_ent.assign((size_t)sym.roww * sym.m_used, 0xFFFFFFFFu);
        sym.row_len.assign((size_t)sym.m_used, 0);
        for (int i = 0; i < sym.m_used; ++i) {
            sym.row_len[(size_t)i] = (uint8_t)ent[(size_t)i].size();
            for (size_t j = 0; j < ent[(size_t)i].size(); ++j) sym.row_ent[j * sym.m_used + i] = ent[(size_t)i][j];
        }
        sym.row_ptr.assign((size_t)sym.m_used + 1, 0);
        sym.row_csr.clear();
        sym.row_range.assign((size_t)sym.m_used, 0);
        for (int i = 0; i < sym.m_used; ++i) {
            sym.row_range[(size_t)i] = ((uint32_t)sym.row_csr.size() << 8) | (uint32_t)ent[(size_t)i].size();
            for (uint32_t e : ent[(size_t)i]) sym.row_csr.push_back(e);
            sym.row_ptr[(size_t)i + 1] = (uint16_t)sym.row_csr.size();
        }
        sym.csc_row.assign((size_t)sym.nnzA, 0);
        for (int64_t k = 0; k < sym.nnzA; ++k) sym.csc_row[(size_t)k] = (uint16_t)sym.row_compact[(size_t)Ai[k]];
        sym.col_start.assign((size_t)sym.n_pad, 0);
        sym.col_len.assign((size_t)sym.n_pad, 0);
        for (int64_t j = 0; j < n; ++j) {
            int64_t len = Ap[j + 1] - Ap[j];
            if (len > 255) return sym.error = "a column of A has more than 255 entries", false;
            sym.colw = std::max(sym.colw, (int)len);
        }
        sym.col_row.assign((size_t)std::max(1, sym.colw) * sym.n_pad, 0);
        sym.col_range.assign((size_t)sym.n_pad, 0);
        for (int64_t j = 0; j < n; ++j)
            sym.col_range[(size_t)sym.var_pos[(size_t)j]] = ((uint32_t)Ap[j] << 8) | (uint32_t)(Ap[j + 1] - Ap[j]);
        for (int64_t j = 0; j < n; ++j) {
            int p = sym.var_pos[(size_t)j];
            sym.col_start[(size_t)p] = (uint32_t)Ap[j];
            sym.col_len[(size_t)p] = (uint8_t)(Ap[j + 1] - Ap[j]);
            for (int64_t k = Ap[j]; k < Ap[j + 1]; ++k)
                sym.col_row[(size_t)(k - Ap[j]) * sym.n_pad + p] = (uint16_t)sym.row_compact[(size_t)Ai[k]];
        }
    }

    // ---- views of P ---------------------------------------------------------------------------------
    {
        sym.Pu_val.resize(pu.size());
        sym.Pu_ij.resize(pu.size());
        std::vector<std::vector<uint32_t>> full((size_t)sym.n_pad);
        for (size_t k = 0; k < pu.size(); ++k) {
            uint32_t pi = (uint32_t)sym.var_pos[(size_t)pu[k].i], pj = (uint32_t)sym.var_pos[(size_t)pu[k].j];
            sym.Pu_val[k] = pu[k].v;
            sym.Pu_ij[k] = (pi << 16) | pj;
            full[pj].push_back((pi << 16) | (uint32_t)k);
            if (pi != pj) full[pi].push_back((pj << 16) | (uint32_t)k);
        }
        for (auto &f : full) sym.pfw = std::max(sym.pfw, (int)f.size());
        if (sym.pfw > 255) return sym.error = "a column of P has more than 255 entries", false;
        sym.pf_ent.assign((size_t)std::max(1, sym.pfw) * sym.n_pad, 0xFFFFFFFFu);
        sym.pf_len.assign((size_t)sym.n_pad, 0);
        sym.pf_ptr.assign((size_t)sym.n_pad + 1, 0);
        sym.pf_csr.clear();
        for (int p = 0; p < sym.n_pad; ++p) {
            for (uint32_t e : full[(size_t)p]) sym.pf_csr.push_back(e);
            sym.pf_ptr[(size_t)p + 1] = (uint16_t)sym.pf_csr.size();
        }
        for (int p = 0; p < sym.n_pad; ++p) {
            sym.pf_len[(size_t)p] = (uint8_t)full[(size_t)p].size();
            for (size_t j = 0; j < full[(size_t)p].size(); ++j) sym.pf_ent[j * sym.n_pad + p] = full[(size_t)p][j];
        }
    }

    // ---- S structure and assembly lists ----------------------------------------------------------------
    const int T = sym.T;
    // off-diagonal pattern: per off block t, per row r (of block t+1): sorted set of columns (of block t)
    std::vector<std::vector<int8_t>> offcols((size_t)std::max(0, T - 1) * bs);
    auto classify = [&](int pa, int pb, int &t, int &r, int &c, bool &diag) -> bool {
        int ta = pa / bs, tb = pb / bs;
        if (ta == tb) {
            diag = true;
            t = ta;
            r = std::max(pa, pb) - ta * bs;
            c = std::min(pa, pb) - ta * bs;
            return true;
        }
        if (std::abs(ta - tb) != 1) return false;
        diag = false;
        t = std::min(ta, tb);
        int hi = ta > tb ? pa : pb, lo = ta > tb ? pb : pa;
        r = hi - (t + 1) * bs;
        c = lo - t * bs;
        return true;
    };
    auto not_tridiagonal = [&]() {
        sym.error = "P + A'A is not block tridiagonal with block size " + std::to_string(bs) + " under the given variable order";
        return false;
    };
    // pass 1: collect off-diagonal pattern
    auto note_off = [&](int t, int r, int c) {
        auto &v = offcols[(size_t)t * bs + r];
        if (std::find(v.begin(), v.end(), (int8_t)c) == v.end()) v.push_back((int8_t)c);
    };
    for (size_t k = 0; k < pu.size(); ++k) {
        int t, r, c;
        bool d;
        if (!classify(sym.var_pos[(size_t)pu[k].i], sym.var_pos[(size_t)pu[k].j], t, r, c, d)) return not_tridiagonal();
        if (!d) note_off(t, r, c);
    }
    std::vector<std::vector<std::pair<int, int>>> row_cells((size_t)sym.m_used); // (pos, k)
    for (int64_t j = 0; j < n; ++j)
        for (int64_t k = Ap[j]; k < Ap[j + 1]; ++k)
            row_cells[(size_t)sym.row_compact[(size_t)Ai[k]]].emplace_back(sym.var_pos[(size_t)j], (int)k);
    for (auto &cells : row_cells) {
        std::sort(cells.begin(), cells.end());
        for (size_t a = 0; a < cells.size(); ++a)
            for (size_t b = a + 1; b < cells.size(); ++b) {
                int t, r, c;
                bool d;
                if (!classify(cells[a].first, cells[b].first, t, r, c, d)) return not_tridiagonal();
                if (!d) note_off(t, r, c);
            }
    }
    for (auto &v : offcols) {
        std::sort(v.begin(), v.end());
        sym.ow = std::max(sym.ow, (int)v.size());
    }
    sym.ow = std::max(sym.ow, 1);
    const int ow = sym.ow;
    const size_t nslots = (size_t)std::max(0, T - 1) * bs * ow;
    sym.off_col.assign(std::max<size_t>(nslots, 1), -1);
    for (size_t tr = 0; tr < offcols.size(); ++tr)
        for (size_t s = 0; s < offcols[tr].size(); ++s) sym.off_col[tr * ow + s] = offcols[tr][s];
    // transposed view
    {
        std::vector<std::vector<std::pair<int8_t, uint16_t>>> tv((size_t)std::max(0, T - 1) * bs);
        for (int t = 0; t + 1 < T; ++t)
            for (int r = 0; r < bs; ++r)
                for (int s = 0; s < ow; ++s) {
                    int c = sym.off_col[((size_t)t * bs + r) * ow + s];
                    if (c >= 0) tv[(size_t)t * bs + c].emplace_back((int8_t)r, (uint16_t)(r * ow + s));
                }
        for (auto &v : tv) sym.owT = std::max(sym.owT, (int)v.size());
        sym.owT = std::max(sym.owT, 1);
        sym.offT_row.assign(std::max<size_t>((size_t)std::max(0, T - 1) * bs * sym.owT, 1), -1);
        sym.offT_slot.assign(sym.offT_row.size(), 0);
        for (size_t tc = 0; tc < tv.size(); ++tc)
            for (size_t s = 0; s < tv[tc].size(); ++s) {
                sym.offT_row[tc * sym.owT + s] = tv[tc][s].first;
                sym.offT_slot[tc * sym.owT + s] = tv[tc][s].second;
            }
    }
    // pass 2: contribution lists
    const size_t ndiag = (size_t)T * sym.dpb;
    std::vector<std::vector<Contrib>> dcon(ndiag), ocon(std::max<size_t>(nslots, 1));
    sym.sd_pk.assign(ndiag, 0xFFFF);
    sym.so_pk.assign(std::max<size_t>(nslots, 1), 0xFFFF);
    auto slot_of = [&](int t, int r, int c) -> size_t {
        for (int s = 0; s < ow; ++s)
            if (sym.off_col[((size_t)t * bs + r) * ow + s] == c) return ((size_t)t * bs + r) * ow + s;
        return (size_t)-1;
    };
    for (size_t k = 0; k < pu.size(); ++k) {
        int t, r, c;
        bool d;
        classify(sym.var_pos[(size_t)pu[k].i], sym.var_pos[(size_t)pu[k].j], t, r, c, d);
        uint16_t &dst = d ? sym.sd_pk[(size_t)t * sym.dpb + r * (r + 1) / 2 + c] : sym.so_pk[slot_of(t, r, c)];
        if (dst != 0xFFFF) return sym.error = "duplicate entry in P", false;
        dst = (uint16_t)k;
    }
    for (int i = 0; i < sym.m_used; ++i) {
        const auto &cells = row_cells[(size_t)i];
        for (size_t a = 0; a < cells.size(); ++a)
            for (size_t b = a; b < cells.size(); ++b) {
                int t, r, c;
                bool d;
                classify(cells[a].first, cells[b].first, t, r, c, d);
                Contrib con{(uint16_t)i, (uint16_t)cells[a].second, (uint16_t)cells[b].second};
                if (d)
                    dcon[(size_t)t * sym.dpb + r * (r + 1) / 2 + c].push_back(con);
                else
                    ocon[slot_of(t, r, c)].push_back(con);
            }
    }
    auto flatten = [](const std::vector<std::vector<Contrib>> &lists, std::vector<uint32_t> &ptr, std::vector<Contrib> &out) {
        ptr.assign(lists.size() + 1, 0);
        for (size_t e = 0; e < lists.size(); ++e) ptr[e + 1] = ptr[e] + (uint32_t)lists[e].size();
        out.clear();
        out.reserve(ptr.back());
        for (const auto &l : lists) out.insert(out.end(), l.begin(), l.end());
    };
    flatten(dcon, sym.sd_ptr, sym.sd_con);
    flatten(ocon, sym.so_ptr, sym.so_con);
    return true;
}

} // namespace gomp
#endif // GOMP_SYMBOLIC_H
