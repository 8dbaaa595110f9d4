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
        if (row_cnt[(size_t)i] > 0) sym.used_rows.push_back((int32_t)i);
        else sym.empty_rows.push_back((int32_t)i);
    }
    // compact order: longest rows first (stable), so that rows a kernel processes side by side have similar lengths
    std::stable_sort(sym.used_rows.begin(), sym.used_rows.end(), [&](int32_t a, int32_t b) { return row_cnt[(size_t)a] > row_cnt[(size_t)b]; });
    for (size_t c = 0; c < sym.used_rows.size(); ++c) sym.row_compact[(size_t)sym.used_rows[c]] = (int32_t)c;
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

// ---- owner-thread ELL views (gomp_duo.cu) ---------------------------------------------------------------
// The two-slot kernel gives every constraint row and every variable position a fixed owner thread out of
// `owners`.  Rows: compact row i belongs to thread i % owners, slot j = i / owners.  Columns: positions are
// sorted by descending entry count and dealt out the same way (col_pos), so that the columns one slot holds
// have similar lengths.  For every slot the entries of all its rows / columns are stored "ELL by slot":
// entry e of the row thread `oid` owns in slot j sits at row_ell[(rbase[j] + e) * owners + oid] -- one
// coalesced, conflict-free load per step, uniform trip counts (rlen[j]) for the whole CTA, no per-row ranges.
// Surplus entries point at zero elements the kernel keeps behind its arrays (value index nnzA / nnzPu,
// position n_pad, row m_used).  Both halves of an entry are stored as BYTE offsets (index * 8), so the kernel adds
// them to a base address without shifting: needs nnzA, m_used, n_pad <= 8190.
struct OwnerEll {
    int owners = 0, rpt = 0, cpt = 0;
    std::vector<int> rlen, rbase, clen, cbase, plen, pbase;
    std::vector<uint32_t> row_ell; // (8 pos << 16 | 8 k)
    std::vector<uint32_t> col_ell; // (8 compact row << 16 | 8 k)
    std::vector<uint32_t> pf_ell;  // (8 other pos << 16 | 8 kPu), full-symmetric view of P
    std::vector<uint16_t> col_pos; // [cpt * owners] position owned by (slot c, thread oid), 0xFFFF = none
};

inline void build_owner_ell(const Symbolic &s, int owners, OwnerEll &o) {
    o = OwnerEll();
    o.owners = owners;
    o.rpt = std::max(1, (s.m_used + owners - 1) / owners);
    o.cpt = std::max(1, (s.n_pad + owners - 1) / owners);
    // rows
    int base = 0;
    for (int j = 0; j < o.rpt; ++j) {
        int len = 0;
        for (int i = j * owners; i < std::min(s.m_used, (j + 1) * owners); ++i) len = std::max(len, (int)(s.row_range[(size_t)i] & 0xFFu));
        o.rlen.push_back(len);
        o.rbase.push_back(base);
        base += len;
    }
    auto pk = [](uint32_t hi, uint32_t lo) { return ((hi * 8u) << 16) | (lo * 8u); };
    o.row_ell.assign((size_t)std::max(1, base) * owners, pk((uint32_t)s.n_pad, (uint32_t)s.nnzA));
    for (int i = 0; i < s.m_used; ++i) {
        const int j = i / owners, oid = i % owners;
        const uint32_t rr = s.row_range[(size_t)i];
        for (uint32_t e = 0; e < (rr & 0xFFu); ++e) {
            const uint32_t en = s.row_csr[(rr >> 8) + e];
            o.row_ell[(size_t)(o.rbase[(size_t)j] + (int)e) * owners + oid] = pk(en >> 16, en & 0xFFFFu);
        }
    }
    // columns: sort positions by (entries of A, entries of P) descending
    std::vector<int> order((size_t)s.n_pad);
    for (int p = 0; p < s.n_pad; ++p) order[(size_t)p] = p;
    auto alen = [&](int p) { return (int)(s.col_range[(size_t)p] & 0xFFu); };
    auto plen = [&](int p) { return (int)s.pf_ptr[(size_t)p + 1] - (int)s.pf_ptr[(size_t)p]; };
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        if (alen(a) != alen(b)) return alen(a) > alen(b);
        return plen(a) > plen(b);
    });
    o.col_pos.assign((size_t)o.cpt * owners, 0xFFFF);
    int cb = 0, pb = 0;
    for (int c = 0; c < o.cpt; ++c) {
        int la = 0, lp = 0;
        for (int q = c * owners; q < std::min(s.n_pad, (c + 1) * owners); ++q) {
            o.col_pos[(size_t)q] = (uint16_t)order[(size_t)q];
            la = std::max(la, alen(order[(size_t)q]));
            lp = std::max(lp, plen(order[(size_t)q]));
        }
        o.clen.push_back(la);
        o.cbase.push_back(cb);
        o.plen.push_back(lp);
        o.pbase.push_back(pb);
        cb += la;
        pb += lp;
    }
    o.col_ell.assign((size_t)std::max(1, cb) * owners, pk((uint32_t)s.m_used, (uint32_t)s.nnzA));
    o.pf_ell.assign((size_t)std::max(1, pb) * owners, pk((uint32_t)s.n_pad, (uint32_t)s.nnzPu));
    for (int q = 0; q < s.n_pad; ++q) {
        const int c = q / owners, oid = q % owners, p = order[(size_t)q];
        const uint32_t cr = s.col_range[(size_t)p];
        for (uint32_t e = 0; e < (cr & 0xFFu); ++e) {
            const uint32_t k = (cr >> 8) + e;
            o.col_ell[(size_t)(o.cbase[(size_t)c] + (int)e) * owners + oid] = pk((uint32_t)s.csc_row[k], k);
        }
        for (int e = 0; e < plen(p); ++e) {
            const uint32_t en = s.pf_csr[(size_t)s.pf_ptr[(size_t)p] + e];
            o.pf_ell[(size_t)(o.pbase[(size_t)c] + e) * owners + oid] = pk(en >> 16, en & 0xFFFFu);
        }
    }
}

// ---- assembly lists "ELL by owner" (gomp_duo.cu) ------------------------------------------------------------
// The two-slot kernel assembles S one factorisation round at a time: round R holds stage R of the down chain and
// stage T-1-R of the up chain (round h = T/2: the middle stage), 2 * dpb packed entries dealt out to the owner
// threads (entry e -> thread e % owners, pass e / owners).  For every (round, pass, thread) the contributions
//   rho[row] * A[k1] * A[k2]
// are stored as 64-bit words (8 row | 8 k1 << 16 | 8 k2 << 32) at con[((round * ept + pass) * maxc + k) * owners + thread]
// -- coalesced loads with a uniform trip count, no pointer chasing; surplus words multiply zeros (row m_used, value
// index nnzA).  meta[(round * ept + pass) * owners + thread] = 8 kPu (or 8 nnzPu: none) | kind << 16 | valid << 18.
// The off-diagonal blocks are assembled once per factorisation the same way (ocon / ometa, entry e of the ELL
// off-diagonal storage -> thread e % owners, pass e / owners).
struct AssemblyEll {
    int owners = 0, rounds = 0, ept = 0, maxc = 0, opt = 0, maxco = 0;
    std::vector<uint64_t> con, ocon;
    std::vector<uint32_t> meta, ometa;
};

inline void build_assembly_ell(const Symbolic &s, int owners, AssemblyEll &a) {
    a = AssemblyEll();
    a.owners = owners;
    const int T = s.T, h = T / 2, dpb = s.dpb;
    a.rounds = h + 1;
    a.ept = (2 * dpb + owners - 1) / owners;
    auto word = [](uint32_t row, uint32_t k1, uint32_t k2) { return (uint64_t)(row * 8u) | ((uint64_t)(k1 * 8u) << 16) | ((uint64_t)(k2 * 8u) << 32); };
    const uint64_t padw = word((uint32_t)s.m_used, (uint32_t)s.nnzA, (uint32_t)s.nnzA);
    auto entry_of = [&](int R, int e, int &ge) { // global packed entry of local entry e of round R, false if none
        const int nb = R == h ? 1 : (R < T - 1 - h ? 2 : 1);
        if (e >= nb * dpb) return false;
        const int which = e >= dpb ? 1 : 0, le = e - which * dpb;
        ge = (which == 0 ? R : T - 1 - R) * dpb + le;
        return true;
    };
    for (int R = 0; R < a.rounds; ++R)
        for (int e = 0; e < a.ept * owners; ++e) {
            int ge;
            if (entry_of(R, e, ge)) a.maxc = std::max(a.maxc, (int)(s.sd_ptr[(size_t)ge + 1] - s.sd_ptr[(size_t)ge]));
        }
    a.maxc = std::max(a.maxc, 1);
    a.con.assign((size_t)a.rounds * a.ept * a.maxc * owners, padw);
    a.meta.assign((size_t)a.rounds * a.ept * owners, (uint32_t)s.nnzPu * 8u);
    for (int R = 0; R < a.rounds; ++R)
        for (int w = 0; w < a.ept; ++w)
            for (int oid = 0; oid < owners; ++oid) {
                int ge;
                if (!entry_of(R, oid + w * owners, ge)) continue;
                const size_t base = ((size_t)(R * a.ept + w) * a.maxc) * owners + oid;
                uint32_t k = 0;
                for (uint32_t q = s.sd_ptr[(size_t)ge]; q < s.sd_ptr[(size_t)ge + 1]; ++q, ++k) {
                    const Contrib &c = s.sd_con[q];
                    a.con[base + (size_t)k * owners] = word(c.row, c.k1, c.k2);
                }
                const int t = ge / dpb, le = ge - t * dpb;
                int r = 0;
                while ((r + 1) * (r + 2) / 2 <= le) ++r;
                const int cc = le - r * (r + 1) / 2;
                const uint32_t kind = r == cc ? (s.pos_var[(size_t)t * s.bs + r] < 0 ? 2u : 1u) : 0u;
                const uint32_t pk = s.sd_pk[(size_t)ge] == 0xFFFF ? (uint32_t)s.nnzPu : (uint32_t)s.sd_pk[(size_t)ge];
                a.meta[(size_t)(R * a.ept + w) * owners + oid] = (pk * 8u) | (kind << 16) | (1u << 18);
            }
    // off-diagonal blocks
    const int noff = std::max(0, T - 1) * s.bs * s.ow;
    const bool has_off = !s.so_ptr.empty() && s.so_ptr.size() > (size_t)noff;
    a.opt = std::max(1, (noff + owners - 1) / owners);
    for (int e = 0; e < noff && has_off; ++e) a.maxco = std::max(a.maxco, (int)(s.so_ptr[(size_t)e + 1] - s.so_ptr[(size_t)e]));
    a.maxco = std::max(a.maxco, 1);
    a.ocon.assign((size_t)a.opt * a.maxco * owners, padw);
    a.ometa.assign((size_t)a.opt * owners, (uint32_t)s.nnzPu * 8u);
    for (int e = 0; e < noff && has_off; ++e) {
        const int w = e / owners, oid = e % owners;
        uint32_t k = 0;
        for (uint32_t q = s.so_ptr[(size_t)e]; q < s.so_ptr[(size_t)e + 1]; ++q, ++k) {
            const Contrib &c = s.so_con[q];
            a.ocon[((size_t)w * a.maxco + k) * owners + oid] = word(c.row, c.k1, c.k2);
        }
        const uint32_t pk = s.so_pk[(size_t)e] == 0xFFFF ? (uint32_t)s.nnzPu : (uint32_t)s.so_pk[(size_t)e];
        a.ometa[(size_t)w * owners + oid] = (pk * 8u) | (1u << 18);
    }
}

} // namespace gomp
#endif // GOMP_SYMBOLIC_H
