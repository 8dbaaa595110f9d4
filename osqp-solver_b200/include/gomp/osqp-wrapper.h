// osqp-wrapper.h -- QPSolver (one QP) and BatchQPSolver (many QPs, one pattern) on libgomp_cuda.
//
// Drop-in for reference src/osqp-wrapper.h:12-60: same class name, same four members, same
// argument meaning and error behaviour --
//   QPSolver(c, P)      builds the problem with q = 0 and OSQP's default settings   (:16-31)
//   update(c)           new A values (same pattern) and bounds; throws std::invalid_argument (:33-43)
//   setWarmStart(x)     primal warm start only                                       (:45-49)
//   solve()             returns {exit code, copy of the primal solution}; iterates, rho and the
//                       factorisation state persist between calls                    (:51-54)
// -- but the work happens in the batched CUDA library behind include/gomp_cuda.h instead of
// osqp-cpp / OSQP / QDLDL.  A QPSolver is a batch of one: correct, but the device is built for
// thousands of QPs at a time -- use BatchQPSolver (or GOMPBatchSolver) for throughput.
// Differences that are deliberate: nothing is printed (the reference prints dimensions and the
// warm-start status and runs OSQP verbose); adaptive_rho_interval is pinned.
#ifndef QP_SOLVER_H
#define QP_SOLVER_H

#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "gomp_cuda.h"
#include "constraints/constraint-builder.h"

namespace gomp {

inline void throwOnError(int rc, const char *what) {
    if (rc != GOMP_OK) throw std::invalid_argument(std::string(what) + ": " + gomp_last_error_string());
}

// Shared pattern + per-problem values of a batch, as plain arrays (what the C ABI consumes).
struct BatchData {
    long long batch = 0, n = 0, m = 0;
    std::vector<long long> A_colptr, A_rowidx;
    std::vector<double> A_val, l, u; // [batch * nnz], [batch * m], [batch * m]
};

} // namespace gomp

class BatchQPSolver {
public:
    // P: shared objective (full symmetric or upper triangular CSC).  data: shared A pattern + values.
    // var_order / block_size: optional block-tridiagonal ordering hint (GOMP: stage-interleaved, 2N).
    BatchQPSolver(const gomp::BatchData &data, const QPMatrixSparse &P, const std::vector<long long> &var_order = {},
                  int block_size = 0, const gomp_settings *settings = nullptr)
        : batch_(data.batch), n_(data.n), m_(data.m), nnz_((long long)data.A_rowidx.size()) {
        gomp_settings st;
        if (settings) st = *settings;
        else gomp_settings_default(&st);
        if (!var_order.empty()) {
            st.var_order = reinterpret_cast<const int64_t *>(var_order.data());
            st.block_size = block_size;
        }
        std::vector<long long> Pp(P.outerIndexPtr(), P.outerIndexPtr() + P.cols() + 1);
        std::vector<long long> Pi(P.innerIndexPtr(), P.innerIndexPtr() + P.nonZeros());
        gomp::throwOnError(gomp_batch_create(&handle_, &st, batch_, n_, m_, reinterpret_cast<const int64_t *>(Pp.data()),
                                             reinterpret_cast<const int64_t *>(Pi.data()), P.valuePtr(),
                                             reinterpret_cast<const int64_t *>(data.A_colptr.data()),
                                             reinterpret_cast<const int64_t *>(data.A_rowidx.data()), data.A_val.data(),
                                             data.l.data(), data.u.data(), nullptr),
                           "gomp_batch_create");
        A_colptr_ = data.A_colptr;
        A_rowidx_ = data.A_rowidx;
    }
    BatchQPSolver(const BatchQPSolver &) = delete;
    BatchQPSolver &operator=(const BatchQPSolver &) = delete;
    ~BatchQPSolver() { gomp_batch_destroy(handle_); }

    // Same pattern required, exactly as osqp-cpp's UpdateConstraintMatrix (reference osqp-wrapper.h:36-38).
    void update(const gomp::BatchData &data) {
        if (data.A_colptr != A_colptr_ || data.A_rowidx != A_rowidx_ || data.batch != batch_)
            throw std::invalid_argument("INVALID_ARGUMENT: constraint matrix sparsity pattern changed");
        gomp::throwOnError(gomp_batch_update(handle_, data.A_val.data(), data.l.data(), data.u.data(), nullptr), "gomp_batch_update");
    }
    void setWarmStart(const std::vector<double> &x0) {
        if ((long long)x0.size() != batch_ * n_) throw std::invalid_argument("INVALID_ARGUMENT: warm start has the wrong size");
        gomp::throwOnError(gomp_batch_warm_start_x(handle_, x0.data()), "gomp_batch_warm_start_x");
    }
    // exit codes (osqp::OsqpExitCode values) and primal solutions [batch * n]
    std::pair<std::vector<ExitCode>, std::vector<double>> solve() {
        gomp::throwOnError(gomp_batch_solve(handle_), "gomp_batch_solve");
        std::vector<double> x((size_t)(batch_ * n_));
        std::vector<int32_t> code((size_t)batch_);
        iters_.resize((size_t)batch_);
        gomp::throwOnError(gomp_batch_get(handle_, x.data(), nullptr, code.data(), iters_.data(), nullptr), "gomp_batch_get");
        std::vector<ExitCode> out((size_t)batch_);
        for (size_t b = 0; b < out.size(); ++b) out[b] = static_cast<ExitCode>(code[b]);
        return {out, x};
    }
    const std::vector<int32_t> &iterations() const { return iters_; }
    long long batch() const { return batch_; }
    gomp_batch *handle() { return handle_; }

private:
    gomp_batch *handle_ = nullptr;
    long long batch_, n_, m_, nnz_;
    std::vector<long long> A_colptr_, A_rowidx_;
    std::vector<int32_t> iters_;
};

class QPSolver {
public:
    QPSolver(const QPConstraints &c, const QPMatrixSparse &P, const std::vector<long long> &var_order = {}, int block_size = 0)
        : batch_(toBatch(c), P, var_order, block_size) {}

    void update(const QPConstraints &qp_constraints) { batch_.update(toBatch(qp_constraints)); }

    void setWarmStart(const QPVector &primal_vector) {
        batch_.setWarmStart(std::vector<double>(primal_vector.data(), primal_vector.data() + primal_vector.size()));
    }

    std::pair<ExitCode, QPVector> solve() {
        auto [codes, x] = batch_.solve();
        QPVector sol((long long)x.size());
        for (size_t k = 0; k < x.size(); ++k) sol[(long long)k] = x[k];
        return {codes[0], sol};
    }

    int iterations() const { return batch_.iterations().empty() ? 0 : batch_.iterations()[0]; }

private:
    static gomp::BatchData toBatch(const QPConstraints &c) {
        const auto &[l, A, u] = c;
        gomp::BatchData d;
        d.batch = 1;
        d.n = A.cols();
        d.m = A.rows();
        d.A_colptr.assign(A.outerIndexPtr(), A.outerIndexPtr() + A.cols() + 1);
        d.A_rowidx.assign(A.innerIndexPtr(), A.innerIndexPtr() + A.nonZeros());
        d.A_val.assign(A.valuePtr(), A.valuePtr() + A.nonZeros());
        d.l.assign(l.data(), l.data() + l.size());
        d.u.assign(u.data(), u.data() + u.size());
        return d;
    }

    BatchQPSolver batch_;
};

#endif // QP_SOLVER_H
