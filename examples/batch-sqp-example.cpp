// batch-sqp-example.cpp -- GOMPSolver::run (one horizon, reference src/gomp-solver.h:57-91) for a batch of the
// repo's synthetic 7-DoF problems, SQP rounds on the GPU.  Host C++ over the two C ABIs only.
// Build:  make -C osqp-solver_b200 example      Run:  build/batch-sqp-example [count] [first_id] [max_rounds] [rerun]
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "gomp-batch-solver.h"

int main(int argc, char **argv) {
    const long long count = argc > 1 ? std::atoll(argv[1]) : 256, first = argc > 2 ? std::atoll(argv[2]) : 0;
    const int max_rounds = argc > 3 ? std::atoi(argv[3]) : GOMPBatchSolver::MAX_ITERATIONS;
    gomp_problem_config cfg;
    gomp_problem_default_config(&cfg);
    try {
        GOMPBatchSolver planner(cfg, first, count);
        const auto t0 = std::chrono::steady_clock::now();
        GOMPBatchSolver::Result r = planner.run(max_rounds);
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        long long optimal = 0, max_r = 0;
        for (long long b = 0; b < count; ++b) {
            optimal += r.exit_code[(size_t)b] == GOMPBatchSolver::kOptimal;
            if (r.rounds[(size_t)b] > max_r) max_r = r.rounds[(size_t)b];
        }
        std::printf("problems %lld  kOptimal %lld  kUnknown %lld  qp_solves %lld  max_rounds_used %lld  %.1f ms\n", count, optimal,
                    count - optimal, r.total_qp_solves, max_r, ms);
        std::printf("codes");
        for (long long b = 0; b < count && b < 64; ++b) std::printf(" %d:%d", (int)r.exit_code[(size_t)b], (int)r.rounds[(size_t)b]);
        std::printf("\n");
        if (argc > 4) { // run the same planner again: a run starts over (fresh QPSolver per run in the reference) and must agree
            GOMPBatchSolver::Result again = planner.run(max_rounds);
            std::printf("rerun_equal %d\n", (int)(again.exit_code == r.exit_code && again.rounds == r.rounds && again.total_qp_solves == r.total_qp_solves));
        }
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 2;
    }
}
