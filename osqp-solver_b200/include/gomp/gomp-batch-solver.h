// gomp-batch-solver.h -- GOMPBatchSolver: GOMPSolver::run for one horizon over a whole batch of problems, with the
// SQP round on the GPU.
//
// Every problem runs the loop of reference src/gomp-solver.h:57-91
//     solve -> not kOptimal: give up | isSolutionOK: done | withObstacles(x) -> update -> again   (<= MAX_ITERATIONS)
// but the QPs are solved by libgomp_cuda as one batch, the feasibility re-check and the re-linearisation are the
// device kernels behind include/gomp_problem.h (gomp_device_check / gomp_device_relinearize) writing A_val / l / u
// in place in the solver's own device copies, and problems that have finished are masked out of the following
// solves (gomp_batch_set_active).  Host C++ only: no CUDA headers, everything goes through the two C ABIs.
// The problems are the repo's synthetic GOMP workload (libgomp_host: gomp_problem_generate), because the device
// kernels need a kinematic model they know; arbitrary RobotBall callbacks stay on GOMPSolver<N>.
#ifndef GOMP_BATCH_SOLVER_H
#define GOMP_BATCH_SOLVER_H

#include <stdexcept>
#include <string>
#include <vector>

#include "gomp_cuda.h"
#include "gomp_problem.h"

class GOMPBatchSolver {
public:
    static constexpr int MAX_ITERATIONS = 100; // reference src/gomp-solver.h:10
    static constexpr int32_t kOptimal = GOMP_EXIT_OPTIMAL, kUnknown = GOMP_EXIT_UNKNOWN;

    struct Result {
        std::vector<int32_t> exit_code;    // kOptimal / kUnknown, as GOMPSolver::run reports a horizon
        std::vector<int32_t> qp_exit_code; // status of the last QP each problem solved
        std::vector<int32_t> rounds;       // QP solves used
        std::vector<double> x;             // [count, n] trajectories (NaN where the last QP was infeasible)
        long long total_qp_solves = 0;
    };

    GOMPBatchSolver(const gomp_problem_config &cfg, long long first_id, long long count, const gomp_settings *settings = nullptr)
        : cfg_(cfg), first_id_(first_id), count_(count) {
        int64_t nnzP = 0;
        check(gomp_problem_dims(&cfg_, &n_, &m_, &nnzA_, &nnzP), "gomp_problem_dims");
        std::vector<int64_t> Ap((size_t)n_ + 1), Ai((size_t)nnzA_), Pp((size_t)n_ + 1), Pi((size_t)nnzP), order((size_t)n_);
        std::vector<double> Px((size_t)nnzP);
        check(gomp_problem_pattern(&cfg_, Ap.data(), Ai.data(), Pp.data(), Pi.data(), Px.data(), order.data()), "gomp_problem_pattern");
        std::vector<double> A_val((size_t)(count * nnzA_)), l((size_t)(count * m_)), u((size_t)(count * m_));
        x0_.resize((size_t)(count * n_));
        check(gomp_problem_generate(&cfg_, first_id, count, A_val.data(), l.data(), u.data(), x0_.data(), nullptr, 0),
              "gomp_problem_generate");
        gomp_settings st;
        if (settings) st = *settings;
        else gomp_settings_default(&st);
        st.var_order = order.data();
        st.block_size = 2 * cfg_.n_dim;
        st.borrow_inputs = 0; // the handle keeps its own device copies, which the device kernels then rewrite
        check(gomp_batch_create(&solver_, &st, count, n_, m_, Pp.data(), Pi.data(), Px.data(), Ap.data(), Ai.data(), A_val.data(),
                                l.data(), u.data(), nullptr),
              "gomp_batch_create");
        check(gomp_batch_device_views(solver_, &dA_, &dl_, &du_, &dx_, &stream_), "gomp_batch_device_views");
        // scene + Cartesian row tables for the device kernels
        std::vector<double> lines((size_t)(count * (cfg_.n_obstacles > 0 ? cfg_.n_obstacles : 1) * 5));
        check(gomp_problem_scene(&cfg_, first_id, count, lines.data()), "gomp_problem_scene");
        int64_t n_rows = 0;
        check(gomp_problem_cart_rows(&cfg_, &n_rows, nullptr, nullptr), "gomp_problem_cart_rows");
        std::vector<int32_t> rows((size_t)n_rows), kidx((size_t)(n_rows * cfg_.n_dim));
        check(gomp_problem_cart_rows(&cfg_, &n_rows, rows.data(), kidx.data()), "gomp_problem_cart_rows");
        check(gomp_device_scene_create(&scene_, &cfg_, count, n_, m_, nnzA_, lines.data(), n_rows, rows.data(), kidx.data(), -1),
              "gomp_device_scene_create");
        check(gomp_device_malloc((void **)&d_ok_, count * (int64_t)sizeof(int32_t), -1), "gomp_device_malloc");
        check(gomp_device_malloc((void **)&d_x0_, count * n_ * (int64_t)sizeof(double), -1), "gomp_device_malloc");
        check(gomp_device_memcpy(d_x0_, x0_.data(), count * n_ * (int64_t)sizeof(double), stream_), "gomp_device_memcpy");
    }
    GOMPBatchSolver(const GOMPBatchSolver &) = delete;
    GOMPBatchSolver &operator=(const GOMPBatchSolver &) = delete;
    ~GOMPBatchSolver() {
        gomp_batch_destroy(solver_);
        gomp_device_scene_destroy(scene_);
        gomp_device_free(d_ok_);
        gomp_device_free(d_x0_);
    }

    Result run(int max_rounds = MAX_ITERATIONS) {
        const size_t B = (size_t)count_;
        Result r;
        r.exit_code.assign(B, kUnknown);
        r.qp_exit_code.assign(B, kUnknown);
        r.rounds.assign(B, 0);
        std::vector<int32_t> running(B, 1), code(B), ok(B);
        if (ran_) {
            // A second run starts over like the first: the rounds rewrote A_val / l / u in place and left the iterates, duals
            // and rho of the last round on the device, whereas the reference builds a fresh QPSolver per run
            // (src/gomp-solver.h:61-65).  Regenerate the initial linearisation and cold-start the handle.
            std::vector<double> A_val((size_t)(count_ * nnzA_)), l((size_t)(count_ * m_)), u((size_t)(count_ * m_));
            check(gomp_problem_generate(&cfg_, first_id_, count_, A_val.data(), l.data(), u.data(), nullptr, nullptr, 0), "gomp_problem_generate");
            check(gomp_batch_setup(solver_, A_val.data(), l.data(), u.data(), nullptr), "gomp_batch_setup");
            check(gomp_batch_sync(solver_), "gomp_batch_sync"); // the host arrays above go out of scope
            check(gomp_batch_device_views(solver_, &dA_, &dl_, &du_, &dx_, &stream_), "gomp_batch_device_views");
        }
        ran_ = true;
        check(gomp_batch_warm_start_x(solver_, x0_.data()), "gomp_batch_warm_start_x");
        check(gomp_batch_set_active(solver_, nullptr), "gomp_batch_set_active");
        for (int round = 0; round < max_rounds; ++round) {
            check(gomp_batch_solve(solver_), "gomp_batch_solve");
            check(gomp_batch_get(solver_, nullptr, nullptr, code.data(), nullptr, nullptr), "gomp_batch_get");
            check(gomp_device_check(scene_, dx_, d_ok_, stream_), "gomp_device_check");
            check(gomp_device_memcpy(ok.data(), d_ok_, (int64_t)(B * sizeof(int32_t)), stream_), "gomp_device_memcpy");
            bool any = false;
            for (size_t b = 0; b < B; ++b) {
                if (!running[b]) continue;
                r.total_qp_solves++;
                r.rounds[b]++;
                r.qp_exit_code[b] = code[b];
                if (code[b] != kOptimal) running[b] = 0;          // gomp-solver.h:74-77: no solution for this horizon
                else if (ok[b]) { r.exit_code[b] = kOptimal; running[b] = 0; } // :78-82
                any = any || running[b];
            }
            if (!any) break;
            check(gomp_batch_set_active(solver_, running.data()), "gomp_batch_set_active");
            // :83-87 withObstacles + update for everybody still running, in place in the handle's device copies
            check(gomp_device_relinearize(scene_, dx_, d_x0_, dA_, dl_, du_, stream_), "gomp_device_relinearize");
            check(gomp_batch_update(solver_, dA_, dl_, du_, nullptr), "gomp_batch_update");
        }
        r.x.resize(B * (size_t)n_);
        check(gomp_device_memcpy(r.x.data(), dx_, (int64_t)(r.x.size() * sizeof(double)), stream_), "gomp_device_memcpy");
        return r;
    }

    long long numVariables() const { return n_; }
    long long numRows() const { return m_; }

private:
    static void check(int rc, const char *what) {
        if (rc != 0) throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " + gomp_last_error_string());
    }
    gomp_problem_config cfg_;
    long long first_id_, count_;
    bool ran_ = false;
    int64_t n_ = 0, m_ = 0, nnzA_ = 0;
    gomp_batch *solver_ = nullptr;
    gomp_device_scene *scene_ = nullptr;
    double *dA_ = nullptr, *dl_ = nullptr, *du_ = nullptr, *d_x0_ = nullptr;
    const double *dx_ = nullptr;
    int32_t *d_ok_ = nullptr;
    void *stream_ = nullptr;
    std::vector<double> x0_;
};

#endif // GOMP_BATCH_SOLVER_H
