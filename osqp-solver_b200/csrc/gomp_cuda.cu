// gomp_cuda.cu -- libgomp_cuda: batched OSQP-style ADMM for GOMP trajectory QPs on B200 (sm_100a).
//
// One persistent CTA per resident slot pulls QPs from a device work queue and runs the whole
// OSQP algorithm for that QP on chip, in fp64:
//   * A values of the QP are staged into shared memory with one TMA bulk copy
//     (cp.async.bulk + mbarrier); the objective P, the ADMM exchange vectors and the factor of
//     the reduced KKT matrix live in shared memory too;
//   * the ADMM iterates that only their owning thread touches (z, y, l, u per constraint row;
//     x, D, q per variable) live in REGISTERS for the whole solve;
//   * Ruiz equilibration, rho classification, assembly of S = P + sigma I + A' diag(rho) A,
//     its block-tridiagonal factorisation (two-way / "twisted": two warps eliminate from both
//     ends towards the middle stage, explicit inverses of the pivot blocks so that every
//     substitution step is one small dense GEMV on warp shuffles), the reduced-form ADMM
//     iteration with fused SpMV + projection + dual update, OSQP's termination /
//     infeasibility tests and adaptive-rho refactorisation all run inside the kernel.
// What it replaces in the reference: everything below the QPSolver seam
// (reference src/osqp-wrapper.h:16-54 -> osqp-cpp -> OSQP osqp_setup / osqp_solve /
// osqp_update_A / osqp_update_bounds / osqp_warm_start_x with QDLDL + AMD).
// The algorithm restated here is checked against oracle/osqp_ref.c in tests/.
//
// No tensor cores: nothing in the loop is a dense contraction larger than 2N x 2N.
// No CPU fallback: every compute entry point fails without a CUDA device.
#include "gomp_cuda.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "gomp_kernel.cuh"
#include "gomp_duo.h"
#include "gomp_stream.h"
#include "symbolic.h"

using namespace gomp_dev;

// kernel variants, each compiled in its own translation unit (gomp_kernel_inst.cu)
#define GOMP_DECLARE_VARIANT(R, C, BSZ)                                                              \
    extern template cudaError_t gomp_dev::prepare_variant<R, C, BSZ>(size_t, int *);                 \
    extern template cudaError_t gomp_dev::launch_variant<R, C, BSZ>(const DevSym &, const DevBatch &, const DevSettings &, int, size_t, cudaStream_t);
GOMP_DECLARE_VARIANT(2, 1, 0)
GOMP_DECLARE_VARIANT(4, 2, 14)
GOMP_DECLARE_VARIANT(4, 2, 0)
GOMP_DECLARE_VARIANT(8, 4, 0)

namespace {

thread_local std::string g_error;
int fail(int code, const std::string &msg) {
    g_error = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return fail(GOMP_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));        \
    } while (0)

// ================================================================================================
//                                           host side
// ================================================================================================
// Device allocations of the handles go through a small per-process cache of freed blocks (per device, keyed by rounded size):
// a planner in the reference's style builds and drops a QPSolver per horizon (src/gomp-solver.h:61-65, ten per plan), and
// cudaMalloc / cudaFree stalls of up to 100 ms each dominated such a plan (5 ms per solve, 14 ms per handle otherwise).  A handle
// waits for its stream before it releases anything, so a cached block is idle.  Blocks above the cap go back to the driver.
class DevCache {
public:
    static DevCache &get() {
        static DevCache c;
        return c;
    }
    cudaError_t alloc(void **p, size_t bytes) {
        const size_t sz = rounded(bytes);
        int dev = 0;
        cudaGetDevice(&dev);
        {
            std::lock_guard<std::mutex> lock(mu_);
            auto it = free_.find({dev, sz});
            if (it != free_.end() && !it->second.empty()) {
                *p = it->second.back();
                it->second.pop_back();
                cached_ -= sz;
                live_[*p] = {dev, sz};
                return cudaSuccess;
            }
        }
        cudaError_t e = cudaMalloc(p, sz);
        if (e != cudaSuccess) { // memory held by the cache may be what is missing: give it back and retry once
            cudaGetLastError();
            trim(0);
            e = cudaMalloc(p, sz);
        }
        if (e == cudaSuccess) {
            std::lock_guard<std::mutex> lock(mu_);
            live_[*p] = {dev, sz};
        }
        return e;
    }
    void release(void *p) {
        if (!p) return;
        std::pair<int, size_t> key;
        {
            std::lock_guard<std::mutex> lock(mu_);
            auto it = live_.find(p);
            if (it == live_.end()) { // not ours
                cudaFree(p);
                return;
            }
            key = it->second;
            live_.erase(it);
            if (cached_ + key.second <= kCap) {
                free_[key].push_back(p);
                cached_ += key.second;
                return;
            }
        }
        int cur = 0;
        cudaGetDevice(&cur);
        if (cur != key.first) cudaSetDevice(key.first);
        cudaFree(p);
        if (cur != key.first) cudaSetDevice(cur);
    }
    void trim(size_t keep) {
        std::vector<std::pair<int, void *>> drop;
        {
            std::lock_guard<std::mutex> lock(mu_);
            for (auto &kv : free_)
                while (!kv.second.empty() && cached_ > keep) {
                    drop.push_back({kv.first.first, kv.second.back()});
                    kv.second.pop_back();
                    cached_ -= kv.first.second;
                }
        }
        int cur = 0;
        cudaGetDevice(&cur);
        for (auto &d : drop) {
            cudaSetDevice(d.first);
            cudaFree(d.second);
        }
        cudaSetDevice(cur);
    }

private:
    static constexpr size_t kCap = (size_t)2 << 30; // bytes kept for reuse per process
    static size_t rounded(size_t bytes) {
        bytes = std::max<size_t>(bytes, 256);
        if (bytes <= ((size_t)1 << 20)) { // next power of two
            size_t s = 256;
            while (s < bytes) s <<= 1;
            return s;
        }
        return (bytes + ((size_t)1 << 20) - 1) & ~(((size_t)1 << 20) - 1); // whole MiB
    }
    std::mutex mu_;
    std::map<std::pair<int, size_t>, std::vector<void *>> free_;
    std::map<void *, std::pair<int, size_t>> live_;
    size_t cached_ = 0;
};

template <typename T>
struct DevArray {
    T *p = nullptr;
    size_t count = 0;
    cudaError_t upload(const std::vector<T> &h) {
        count = h.size();
        cudaError_t e = DevCache::get().alloc((void **)&p, std::max<size_t>(count, 1) * sizeof(T));
        if (e != cudaSuccess) return e;
        if (count) e = cudaMemcpy(p, h.data(), count * sizeof(T), cudaMemcpyHostToDevice);
        return e;
    }
    cudaError_t alloc(size_t n) {
        count = n;
        return DevCache::get().alloc((void **)&p, std::max<size_t>(n, 1) * sizeof(T));
    }
    void release() {
        if (p) DevCache::get().release(p);
        p = nullptr;
    }
};

bool is_device_pointer(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

using PrepareFn = cudaError_t (*)(size_t, int *);
using LaunchFn = cudaError_t (*)(const DevSym &, const DevBatch &, const DevSettings &, int, size_t, cudaStream_t);
struct KernelChoice {
    PrepareFn prepare = nullptr;
    LaunchFn launch = nullptr;
    int rpt = 0, cpt = 0, bs = 0;
};
// smallest register-resident variant that covers (rows, columns) per thread; the compile-time
// block size is used when it matches, otherwise the generic (runtime block size) variant
KernelChoice pick_kernel(int m_used, int n_pad, int bs) {
    KernelChoice k;
    const int need_r = std::max(1, (m_used + kOwners - 1) / kOwners), need_c = (n_pad + kOwners - 1) / kOwners;
    struct Opt { int r, c, b; PrepareFn p; LaunchFn l; };
    const Opt opts[] = {
        {2, 1, 0, prepare_variant<2, 1, 0>, launch_variant<2, 1, 0>},
        {4, 2, 14, prepare_variant<4, 2, 14>, launch_variant<4, 2, 14>},
        {4, 2, 0, prepare_variant<4, 2, 0>, launch_variant<4, 2, 0>},
        {8, 4, 0, prepare_variant<8, 4, 0>, launch_variant<8, 4, 0>},
    };
    for (const Opt &o : opts)
        if (need_r <= o.r && need_c <= o.c && (o.b == 0 || o.b == bs)) {
            k.prepare = o.p; k.launch = o.l; k.rpt = o.r; k.cpt = o.c; k.bs = o.b;
            return k;
        }
    return k;
}

// ---- roofline denominators measured on the device this library runs on ------------------------------
// MEASURED_PEAKS.json only carries HBM and bf16 tensor numbers; the kernel here is bounded by the FP64
// pipe and by shared-memory bandwidth, so both are measured with two tiny kernels.
__global__ void peak_fp64_kernel(double *out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0, a4 = a0 + 4.0, a5 = a0 + 5.0, a6 = a0 + 6.0, a7 = a0 + 7.0;
    const double m = 1.0000001, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
// Work-queue order for a warm re-solve: problems in descending order of the iterations their previous solve took
// (bucketed), so that the few long solves of a batch start first instead of forming its tail.  One CTA.
// exit_code_of maps OSQP's status_val to osqp-cpp's OsqpExitCode; exit_code_kernel does so on the device so
// that results can be handed to a device consumer without a host round trip.
__host__ __device__ __forceinline__ int exit_code_of(int status) {
    switch (status) {
    case 1: return GOMP_EXIT_OPTIMAL;
    case 2: return GOMP_EXIT_OPTIMAL_INACCURATE;
    case -3: return GOMP_EXIT_PRIMAL_INFEASIBLE;
    case 3: return GOMP_EXIT_PRIMAL_INFEASIBLE_INACCURATE;
    case -4: return GOMP_EXIT_DUAL_INFEASIBLE;
    case 4: return GOMP_EXIT_DUAL_INFEASIBLE_INACCURATE;
    case -2: return GOMP_EXIT_MAX_ITERATIONS;
    case -7: return GOMP_EXIT_NON_CONVEX;
    default: return GOMP_EXIT_UNKNOWN;
    }
}
__global__ void exit_code_kernel(const int *__restrict__ status, int n, int *__restrict__ code) {
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < n; b += gridDim.x * blockDim.x) code[b] = exit_code_of(status[b]);
}
__global__ void order_by_iters_kernel(const int *__restrict__ iters, int n, int *__restrict__ order) {
    __shared__ int hist[257], base[257];
    for (int k = threadIdx.x; k < 257; k += blockDim.x) hist[k] = 0;
    __syncthreads();
    auto bucket = [](int it) { return 255 - min(255, max(0, it) >> 4); }; // long solves -> low bucket index -> front of the queue
    for (int b = threadIdx.x; b < n; b += blockDim.x) atomicAdd(&hist[bucket(iters[b])], 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int k = 0; k < 256; ++k) {
            base[k] = acc;
            acc += hist[k];
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < n; b += blockDim.x) order[atomicAdd(&base[bucket(iters[b])], 1)] = b;
}

// Infeasibility certificates of the last solve (OSQP check_termination + store_solution: delta_y projected onto the polar of
// the recession cone of [l, u], un-scaled with E and normalised to unit infinity norm; delta_x un-scaled with D likewise).
// The ADMM kernels leave delta_y / delta_x of the terminating iteration in the handle's scratch arrays (compact rows /
// positions, scaled space); this kernel expands them into the caller's row / variable order.  One CTA per QP.
__global__ void certificate_kernel(int m, int n, int mu, int np, const int *__restrict__ used_rows, const int *__restrict__ pos_var,
                                   const int *__restrict__ status, const int *__restrict__ iters, const int *__restrict__ empty_rows, int n_empty,
                                   const double *__restrict__ l, const double *__restrict__ u,
                                   const double *__restrict__ Eg, const double *__restrict__ Dg, const double *__restrict__ dyg,
                                   const double *__restrict__ dxg, int scaled_termination, double *__restrict__ prim, double *__restrict__ dual) {
    __shared__ double red[32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int st = status[b];
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    auto block_max = [&](double v) {
        for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
        __syncthreads();
        if ((tid & 31) == 0) red[tid >> 5] = v;
        __syncthreads();
        double r = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) r = fmax(r, red[k]);
        return r;
    };
    if (prim) {
        double *out = prim + (size_t)b * m;
        const bool has = st == -3 || st == 3;
        for (int i = tid; i < m; i += blockDim.x) out[i] = has ? 0.0 : qnan; // rows without entries never leave y = 0
        __syncthreads();
        if (has && iters[b] == 0) {
            // infeasible before the first iteration: a row of A without entries whose bounds exclude 0 (0 = A_r x is not in
            // [l_r, u_r]); the unit vector of that row, signed towards the violated side, is the certificate
            for (int e = tid; e < n_empty; e += blockDim.x) {
                const int r = empty_rows[e];
                if (l[(size_t)b * m + r] > 0.0) out[r] = -1.0;
                else if (u[(size_t)b * m + r] < 0.0) out[r] = 1.0;
            }
        } else if (has) {
            double nrm = 0.0;
            for (int i = tid; i < mu; i += blockDim.x) {
                const int orig = used_rows[i];
                const double e = Eg[(size_t)b * mu + i];
                const double lo = fmax(l[(size_t)b * m + orig], -kInfty) * e, hi = fmin(u[(size_t)b * m + orig], kInfty) * e;
                double d = dyg[(size_t)b * mu + i];
                if (hi > kInfty * kMinScaling) d = lo < -kInfty * kMinScaling ? 0.0 : fmin(d, 0.0);
                else if (lo < -kInfty * kMinScaling) d = fmax(d, 0.0);
                if (!scaled_termination) d *= e;
                out[orig] = d;
                nrm = fmax(nrm, fabs(d));
            }
            nrm = block_max(nrm);
            const double inv = 1.0 / nrm;
            for (int i = tid; i < mu; i += blockDim.x) out[used_rows[i]] *= inv;
        }
    }
    if (dual) {
        double *out = dual + (size_t)b * n;
        const bool has = (st == -4 || st == 4) && dxg != nullptr;
        for (int j = tid; j < n; j += blockDim.x) out[j] = qnan;
        __syncthreads();
        if (has) {
            double nrm = 0.0;
            for (int p = tid; p < np; p += blockDim.x) {
                const int var = pos_var[p];
                if (var < 0) continue;
                double d = dxg[(size_t)b * np + p];
                if (!scaled_termination) d *= Dg[(size_t)b * np + p];
                out[var] = d;
                nrm = fmax(nrm, fabs(d));
            }
            nrm = block_max(nrm);
            const double inv = 1.0 / nrm;
            for (int p = tid; p < np; p += blockDim.x)
                if (pos_var[p] >= 0) out[pos_var[p]] *= inv;
        }
    }
}

// l <= u for device-resident bounds (the host path checks on the CPU before copying): one flag per handle, read back with
// the results.  OSQP refuses such data (validate_data / osqp_update_bounds), the reference asserts / throws on it.
__global__ void bounds_check_kernel(const double *__restrict__ l, const double *__restrict__ u, size_t total, int *flag) {
    int bad = 0;
    for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (size_t)gridDim.x * blockDim.x) bad |= l[k] > u[k];
    if (__syncthreads_or(bad) && threadIdx.x == 0) atomicOr(flag, 1);
}

// Work-queue order for a COLD solve with a primal warm start (what the reference does once per horizon, src/gomp-solver.h:61-71):
// the QPs whose warm start violates its constraints most go first.  How far x0 is from feasible is the cheapest predictor of how
// long ADMM will take that the data offers -- on the GOMP workload the QPs that run to max_iter rank in the first 6-16 %, the
// infeasible ones (hundreds of iterations) at a median of 14 % -- and a batch ends with its longest solves: starting them first
// instead of wherever their index falls shortens the tail in which a handful of SMs finish alone (longest-processing-time-first
// scheduling with estimated times).  The order changes nothing about any QP's result.  One warp per QP: key = quantised
// log2(1 + 1e4 * max_i dist(x0, {x: l_i <= a_i'x <= u_i})), in the units order_by_iters_kernel buckets.
__global__ void predict_key_kernel(const DevSym S, const double *__restrict__ A_val, long long A_pitch, const double *__restrict__ l,
                                   const double *__restrict__ u, const double *__restrict__ x0, int batch, int *__restrict__ key) {
    const int b = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (b >= batch) return;
    const double *A = A_val + (size_t)b * A_pitch, *xb = x0 + (size_t)b * S.n, *lb = l + (size_t)b * S.m, *ub = u + (size_t)b * S.m;
    double viol = 0.0;
    for (int i = lane; i < S.m_used; i += 32) {
        const uint32_t rg = __ldg(S.row_range + i);
        double ax = 0.0, nn = 0.0;
        for (int q = (int)(rg >> 8), q1 = q + (int)(rg & 0xFFu); q < q1; ++q) {
            const uint32_t e = __ldg(S.row_csr + q);
            const int var = __ldg(S.pos_var + (e >> 16));
            const double a = A[e & 0xFFFFu];
            ax = fma(a, var >= 0 ? xb[var] : 0.0, ax);
            nn = fma(a, a, nn);
        }
        const int orig = __ldg(S.used_rows + i);
        // distance of x0 to the row's slab l <= a'x <= u (Euclidean: violation / |a|): ranks the long solves better than the raw
        // violation (the two max_iter QPs of the bench batch at 47 and 197 of 4096, the 200+ iteration ones at a median of 237)
        viol = fmax(viol, fmax(lb[orig] - ax, ax - ub[orig]) * rsqrt(nn + 1e-300));
    }
    for (int o = 16; o; o >>= 1) viol = fmax(viol, __shfl_xor_sync(0xffffffffu, viol, o));
    if (lane == 0) {
        const double v = viol > 0.0 ? (viol < 1e30 ? viol : 1e30) : 0.0;
        key[b] = 16 * min(255, (int)(15.0 * log2(1.0 + 1e4 * v)));
    }
}

__global__ void peak_smem_kernel(double *out, int iters) {
    extern __shared__ __align__(16) double buf[];
    const int words = 4096; // 32 KB of doubles per CTA
    for (int i = threadIdx.x; i < words; i += blockDim.x) buf[i] = i * 1e-6;
    __syncthreads();
    const double2 *v = reinterpret_cast<const double2 *>(buf);
    double2 acc = make_double2(0.0, 0.0);
    int idx = threadIdx.x;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { // conflict-free 16-byte loads, 8 in flight
            const double2 t = v[(idx + k * 256) & (words / 2 - 1)];
            acc.x += t.x;
            acc.y += t.y;
        }
        idx = (idx + 37) & (words / 2 - 1);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y;
}

} // namespace

struct gomp_batch {
    gomp_settings st;
    gomp::Symbolic sym;
    DevSym ds;
    DevBatch db;
    DevSettings dst;
    KernelChoice kernel;
    bool streamed = false; // long-horizon back end (gomp_stream.cu): QP and factor in a global workspace
    DevStream dx;
    bool duo = false;      // two-slot back end (gomp_duo.cu): two QPs per SM, factor in tensor memory
    DevDuo dd;
    gomp::OwnerEll oell;
    gomp::AssemblyEll aell;
    DevArray<unsigned long long> d_ascon, d_asocon;
    DevArray<uint32_t> d_asmeta, d_asometa;
    DevArray<uint32_t> d_rell, d_cell, d_pell;
    DevArray<uint16_t> d_cpos;
    DevArray<double> d_ws, d_pre;
    bool no_predict = false; // GOMP_NO_PREDICT_ORDER at create: cold solves keep index order in the work queue
    int max_clusters = 0;  // ... and how many such clusters are resident at once
    int solve_cluster = 1; // streamed back end: CTAs per QP of the ADMM kernel itself (helpers share the sparse passes)
    int pre_cluster = 0; // streamed back end: CTAs per QP of the Ruiz cluster kernel that runs before the ADMM kernel (0 = in-kernel Ruiz)
    DevArray<uint32_t> d_row_ell;
    DevArray<uint16_t> d_col_ell;
    int device = 0, num_sms = 0, ctas_per_sm = 0, grid = 0;
    size_t smem_bytes = 0, workspace_bytes = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evc0 = nullptr, evc1 = nullptr;
    int max_ctas = 0;        // gomp_batch_set_max_ctas: cap on the CTAs of the following solves (0 = one per resident slot)
    bool iters_fresh = false; // d_iters / d_status hold the previous solve of the problems now loaded
    bool code_valid = false;  // d_code holds the exit codes of the last solve
    int64_t batch = 0, launches = 0;
    double last_solve_ms = 0, last_h2d_ms = 0;
    bool has_q = false, data_loaded = false;
    int pending_flags = 1; // bit0: cold start, bit1: warm start x pending -- consumed by the next solve
    const double *x0_src = nullptr;
    // pattern arrays on the device
    DevArray<int32_t> d_pos_var, d_used_rows, d_empty_rows;
    DevArray<uint32_t> d_row_csr, d_row_range, d_col_range, d_pf_csr, d_Pu_ij, d_sd_ptr, d_so_ptr;
    DevArray<uint16_t> d_row_ptr, d_csc_row, d_pf_ptr, d_offT_slot, d_sd_pk, d_so_pk, d_tri_rc;
    DevArray<uint8_t> d_sd_kind;
    DevArray<int8_t> d_off_col, d_offT_row;
    DevArray<double> d_Pu_val;
    DevArray<gomp::Contrib> d_sd_con, d_so_con;
    // per-batch storage
    DevArray<double> d_A, d_l, d_u, d_q, d_x0, d_xs, d_zs, d_ys, d_rho, d_E, d_D, d_qs, d_dy, d_dx, d_Lg, d_Ug, d_xo, d_yo, d_pri, d_dua, d_obj;
    DevArray<unsigned long long> d_prof;
    DevArray<int> d_status, d_iters, d_rup, d_counter, d_active, d_order, d_code, d_flag, d_pred;
    int h_flag = 0; // d_flag as of the last wait on the stream: device-resident bounds with l > u
    bool order_valid = false; // d_order holds the order derived from the previous solve of the same problems
    const int *active_src = nullptr; // per-QP mask of the next solves (device), NULL = all
    const double *A_src = nullptr, *l_src = nullptr, *u_src = nullptr, *q_src = nullptr; // borrowed device inputs
    long long A_pitch = 0;
};

namespace {

int upload_pattern(gomp_batch *h) {
    const gomp::Symbolic &s = h->sym;
#define UP(dst, src) CUDA_TRY(h->dst.upload(s.src))
    UP(d_pos_var, pos_var);
    UP(d_used_rows, used_rows);
    UP(d_empty_rows, empty_rows);
    UP(d_row_csr, row_csr);
    UP(d_row_ptr, row_ptr);
    UP(d_row_range, row_range);
    UP(d_csc_row, csc_row);
    UP(d_col_range, col_range);
    UP(d_pf_csr, pf_csr);
    UP(d_pf_ptr, pf_ptr);
    UP(d_Pu_val, Pu_val);
    UP(d_Pu_ij, Pu_ij);
    UP(d_off_col, off_col);
    UP(d_offT_row, offT_row);
    UP(d_offT_slot, offT_slot);
    UP(d_sd_ptr, sd_ptr);
    UP(d_sd_con, sd_con);
    UP(d_sd_pk, sd_pk);
    UP(d_so_ptr, so_ptr);
    UP(d_so_con, so_con);
    UP(d_so_pk, so_pk);
#undef UP
    std::vector<uint8_t> kind((size_t)s.T * s.dpb, 0);
    for (int t = 0; t < s.T; ++t)
        for (int r = 0; r < s.bs; ++r) kind[(size_t)t * s.dpb + r * (r + 1) / 2 + r] = s.pos_var[(size_t)t * s.bs + r] < 0 ? 2 : 1;
    CUDA_TRY(h->d_sd_kind.upload(kind));
    std::vector<uint16_t> tri;
    for (int r = 0; r < 32; ++r)
        for (int c = 0; c <= r; ++c) tri.push_back((uint16_t)((r << 8) | c));
    CUDA_TRY(h->d_tri_rc.upload(tri));

    DevSym &d = h->ds;
    d.n = (int)s.n; d.m = (int)s.m; d.nnzA = (int)s.nnzA; d.nnzPu = (int)s.nnzPu; d.m_used = s.m_used; d.bs = s.bs;
    d.T = s.T; d.n_pad = s.n_pad; d.ow = s.ow; d.owT = s.owT;
    d.n_empty = (int)s.empty_rows.size(); d.noff = std::max(0, s.T - 1) * s.bs * s.ow; d.npf = (int)s.pf_csr.size();
    d.uniform_off = s.T >= 2 ? 1 : 0;
    for (int t = 1; t + 1 < s.T && d.uniform_off; ++t) {
        for (int e = 0; e < s.bs * s.ow; ++e)
            if (s.off_col[(size_t)t * s.bs * s.ow + e] != s.off_col[e]) d.uniform_off = 0;
        for (int e = 0; e < s.bs * s.owT; ++e)
            if (s.offT_row[(size_t)t * s.bs * s.owT + e] != s.offT_row[e] || s.offT_slot[(size_t)t * s.bs * s.owT + e] != s.offT_slot[e])
                d.uniform_off = 0;
    }
    d.pos_var = h->d_pos_var.p; d.used_rows = h->d_used_rows.p; d.empty_rows = h->d_empty_rows.p;
    d.row_csr = h->d_row_csr.p; d.row_range = h->d_row_range.p; d.col_range = h->d_col_range.p; d.pf_csr = h->d_pf_csr.p; d.Pu_ij = h->d_Pu_ij.p;
    d.row_ptr = h->d_row_ptr.p; d.csc_row = h->d_csc_row.p; d.pf_ptr = h->d_pf_ptr.p; d.offT_slot = h->d_offT_slot.p;
    d.off_col = h->d_off_col.p; d.offT_row = h->d_offT_row.p; d.Pu_val = h->d_Pu_val.p;
    d.sd_ptr = h->d_sd_ptr.p; d.sd_con = h->d_sd_con.p; d.sd_pk = h->d_sd_pk.p; d.sd_kind = h->d_sd_kind.p;
    d.so_ptr = h->d_so_ptr.p; d.so_con = h->d_so_con.p; d.so_pk = h->d_so_pk.p; d.tri_rc = h->d_tri_rc.p;
    // shared-memory carve-up (byte offsets; every region 16-byte aligned)
    size_t o = 0;
    auto take = [&o](size_t bytes) {
        size_t at = o;
        o += (bytes + 15) & ~(size_t)15;
        return (int)at;
    };
    const size_t bb = (size_t)d.bs * d.bs, noffT = (size_t)std::max(0, d.T - 1) * d.bs * d.owT;
    d.oA = take(8 * (size_t)d.nnzA);
    d.oP = take(8 * (size_t)std::max(1, d.nnzPu));
    d.oW = take(8 * (size_t)std::max(1, d.m_used));
    d.oRHS = take(8 * (size_t)d.n_pad);
    d.oDinv = take(8 * (size_t)d.T * bb);
    d.oOff = take(8 * (size_t)std::max(1, d.noff));
    d.oV = take(8 * 512); // per-warp broadcast lines (256 bytes apart, double-buffered)
    d.oScr = take(8 * std::max<size_t>(2 * bb, 256)); // per-warp scratch (generic path) / pivot row+column exchange of the factor groups
    d.oRed = take(8 * (size_t)kRedK * (kThreads / 32));
    d.oRowRange = take(4 * (size_t)std::max(1, d.m_used));
    d.oRowEnt = take(4 * ((size_t)d.nnzA + 4));
    d.oColRange = take(4 * (size_t)d.n_pad);
    d.oCscRow = take(2 * ((size_t)d.nnzA + 4));
    d.oOffTslot = take(2 * std::max<size_t>(1, noffT));
    d.oOffCol = take(std::max<size_t>(1, (size_t)d.noff));
    d.oOffTrow = take(std::max<size_t>(1, noffT));
    d.oRowPtr = 0;
    // optional residents, in order of benefit, as long as everything still fits in 227 KB per CTA:
    // the scaled bounds l, u (read every iteration) and the P-side pattern arrays (scaling / residuals only)
    const size_t kSmemMax = (size_t)227 * 1024;
    {
        const size_t before = o;
        d.oLo = take(8 * (size_t)std::max(1, d.m_used));
        d.oHi = take(8 * (size_t)std::max(1, d.m_used));
        d.bounds_in_smem = 1;
        if (o + 16 > kSmemMax) {
            o = before;
            d.bounds_in_smem = 0;
        }
    }
    {
        const size_t before = o;
        d.oPfEnt = take(4 * (size_t)std::max(1, d.npf));
        d.oPuIJ = take(4 * (size_t)std::max(1, d.nnzPu));
        d.oPfPtr = take(2 * ((size_t)d.n_pad + 1));
        d.p_in_smem = 1;
        if (o + 16 > kSmemMax) {
            o = before;
            d.p_in_smem = 0;
        }
    }
    d.oBar = take(16);
    h->smem_bytes = o;
    return GOMP_OK;
}

// l <= u everywhere (host arrays), on a few threads
bool bounds_ordered(const double *l, const double *u, size_t total) {
    const unsigned nt = (unsigned)std::max<size_t>(1, std::min<size_t>(std::min<size_t>(16, std::thread::hardware_concurrency()), total / 65536 + 1));
    std::atomic<int> bad{0};
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nt; ++t)
        pool.emplace_back([&, t]() {
            const size_t a = total * t / nt, b = total * (t + 1) / nt;
            int any = 0;
            for (size_t k = a; k < b; ++k) any |= (l[k] > u[k]);
            if (any) bad.store(1);
        });
    for (auto &th : pool) th.join();
    return bad.load() == 0;
}

// copy (or borrow) one input array; `dst` is allocated lazily
int stage_input(gomp_batch *h, const double *src, DevArray<double> &dst, size_t count, const double **use) {
    if (!src) return GOMP_OK;
    const bool on_device = is_device_pointer(src);
    if (on_device && h->st.borrow_inputs) {
        *use = src;
        return GOMP_OK;
    }
    if (!dst.p) {
        CUDA_TRY(dst.alloc(count));
        h->workspace_bytes += count * sizeof(double);
    }
    if (src == dst.p) { // the caller rewrote the handle's own copy in place (gomp_batch_device_views)
        *use = dst.p;
        return GOMP_OK;
    }
    CUDA_TRY(cudaMemcpyAsync(dst.p, src, count * sizeof(double), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                             h->stream));
    *use = dst.p;
    return GOMP_OK;
}

int load_data(gomp_batch *h, const double *A_val, const double *l, const double *u, const double *q, bool cold) {
    const gomp::Symbolic &s = h->sym;
    const size_t B = (size_t)h->batch;
    CUDA_TRY(cudaSetDevice(h->device));
    if (cold && (!A_val || !l || !u)) return fail(GOMP_ERR_INVALID_ARGUMENT, "setup needs A_val, l and u");
    // bounds validation (OSQP validate_data / osqp_update_bounds: l <= u) on host inputs
    if (l && u && !is_device_pointer(l) && !is_device_pointer(u) && !bounds_ordered(l, u, B * (size_t)s.m))
        return fail(GOMP_ERR_INVALID_ARGUMENT, "lower bound exceeds upper bound");
    CUDA_TRY(cudaEventRecord(h->evc0, h->stream));
    if (l && u && is_device_pointer(l) && is_device_pointer(u)) {
        // device-resident bounds are checked on the device; the verdict surfaces at the next gomp_batch_get / _sync
        if (!h->d_flag.p) CUDA_TRY(h->d_flag.alloc(1));
        CUDA_TRY(cudaMemsetAsync(h->d_flag.p, 0, sizeof(int), h->stream));
        const size_t total = B * (size_t)s.m;
        bounds_check_kernel<<<(unsigned)std::min<size_t>((total + 255) / 256, (size_t)4 * std::max(1, h->num_sms)), 256, 0, h->stream>>>(l, u, total, h->d_flag.p);
        CUDA_TRY(cudaGetLastError());
        h->launches += 1;
    }
    int rc;
    if ((rc = stage_input(h, A_val, h->d_A, B * (size_t)s.nnzA, &h->A_src))) return rc;
    if ((rc = stage_input(h, l, h->d_l, B * (size_t)s.m, &h->l_src))) return rc;
    if ((rc = stage_input(h, u, h->d_u, B * (size_t)s.m, &h->u_src))) return rc;
    if (q) {
        if ((rc = stage_input(h, q, h->d_q, B * (size_t)s.n, &h->q_src))) return rc;
        if (!h->d_qs.p) {
            CUDA_TRY(h->d_qs.alloc(B * (size_t)s.n_pad));
            h->workspace_bytes += B * (size_t)s.n_pad * sizeof(double);
        }
        h->has_q = true;
    }
    CUDA_TRY(cudaEventRecord(h->evc1, h->stream));
    if (cold) {
        h->pending_flags = 1;   // cold start; a pending warm start belongs to the previous problem
        h->order_valid = false; // new problems: the previous iteration counts say nothing about them
        h->iters_fresh = false;
    }
    h->data_loaded = true;
    return GOMP_OK;
}

} // namespace

extern "C" {

int gomp_settings_default(gomp_settings *s) {
    if (!s) return GOMP_ERR_INVALID_ARGUMENT;
    std::memset(s, 0, sizeof *s);
    s->rho = 0.1;
    s->sigma = 1e-6;
    s->alpha = 1.6;
    s->scaling = 10;
    s->adaptive_rho = 1;
    s->adaptive_rho_interval = 25;
    s->adaptive_rho_tolerance = 5.0;
    s->max_iter = 4000;
    s->eps_abs = 1e-3;
    s->eps_rel = 1e-3;
    s->eps_prim_inf = 1e-4;
    s->eps_dual_inf = 1e-4;
    s->scaled_termination = 0;
    s->check_termination = 25;
    s->warm_start = 1;
    s->verbose = 0;
    s->device = -1;
    s->stream = nullptr;
    s->linsys = GOMP_LINSYS_AUTO;
    s->block_size = 0;
    s->var_order = nullptr;
    s->threads_per_cta = 0;
    s->borrow_inputs = 0;
    s->prefer_throughput = 0;
    return GOMP_OK;
}

int gomp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int gomp_measure_peaks(double *fp64_tflops, double *smem_gbs) {
    if (gomp_device_count() <= 0) return fail(GOMP_ERR_NO_DEVICE, "no CUDA device visible");
    cudaDeviceProp prop;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
    const int threads = 256, grid = prop.multiProcessorCount * 8;
    double *out = nullptr;
    CUDA_TRY(cudaMalloc((void **)&out, (size_t)grid * threads * sizeof(double)));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    float ms = 0;
    if (fp64_tflops) {
        const int iters = 1 << 15;
        double best = 0;
        for (int rep = 0; rep < 4; ++rep) {
            CUDA_TRY(cudaEventRecord(e0));
            peak_fp64_kernel<<<grid, threads>>>(out, iters);
            CUDA_TRY(cudaEventRecord(e1));
            CUDA_TRY(cudaEventSynchronize(e1));
            CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
            best = std::max(best, 2.0 * 8.0 * iters * (double)grid * threads / (ms * 1e-3) / 1e12);
        }
        *fp64_tflops = best;
    }
    if (smem_gbs) {
        const int iters = 1 << 13;
        double best = 0;
        for (int rep = 0; rep < 4; ++rep) {
            CUDA_TRY(cudaEventRecord(e0));
            peak_smem_kernel<<<grid, threads, 32768>>>(out, iters);
            CUDA_TRY(cudaEventRecord(e1));
            CUDA_TRY(cudaEventSynchronize(e1));
            CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
            best = std::max(best, 16.0 * 8.0 * iters * (double)grid * threads / (ms * 1e-3) / 1e9);
        }
        *smem_gbs = best;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return GOMP_OK;
}

const char *gomp_version(void) { return "gomp_cuda 0.1 (sm_100a)"; }
const char *gomp_last_error_string(void) { return g_error.c_str(); }

int gomp_batch_destroy(gomp_batch *h) {
    if (!h) return GOMP_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    h->d_pos_var.release(); h->d_used_rows.release(); h->d_empty_rows.release(); h->d_row_csr.release(); h->d_row_range.release(); h->d_col_range.release();
    h->d_pf_csr.release(); h->d_Pu_ij.release(); h->d_sd_ptr.release(); h->d_so_ptr.release(); h->d_row_ptr.release();
    h->d_csc_row.release(); h->d_pf_ptr.release(); h->d_offT_slot.release(); h->d_sd_pk.release(); h->d_so_pk.release();
    h->d_sd_kind.release(); h->d_off_col.release(); h->d_offT_row.release(); h->d_Pu_val.release(); h->d_sd_con.release();
    h->d_so_con.release();
    h->d_A.release(); h->d_l.release(); h->d_u.release(); h->d_q.release(); h->d_x0.release(); h->d_xs.release();
    h->d_zs.release(); h->d_ys.release(); h->d_rho.release(); h->d_E.release(); h->d_D.release(); h->d_qs.release(); h->d_dx.release(); h->d_Lg.release(); h->d_Ug.release(); h->d_tri_rc.release(); h->d_dy.release(); h->d_xo.release();
    h->d_yo.release(); h->d_pri.release(); h->d_dua.release(); h->d_obj.release();
    h->d_rell.release(); h->d_cell.release(); h->d_pell.release(); h->d_cpos.release();
    h->d_ascon.release(); h->d_asocon.release(); h->d_asmeta.release(); h->d_asometa.release();
    h->d_active.release(); h->d_order.release(); h->d_code.release(); h->d_flag.release(); h->d_pre.release(); h->d_pred.release(); h->d_ws.release(); h->d_row_ell.release(); h->d_col_ell.release(); h->d_status.release(); h->d_iters.release(); h->d_rup.release(); h->d_counter.release(); h->d_prof.release();
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->evc0) cudaEventDestroy(h->evc0);
    if (h->evc1) cudaEventDestroy(h->evc1);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return GOMP_OK;
}

int gomp_batch_create(gomp_batch **out, const gomp_settings *settings, int64_t batch, int64_t n, int64_t m,
                      const int64_t *P_colptr, const int64_t *P_rowidx, const double *P_val, const int64_t *A_colptr,
                      const int64_t *A_rowidx, const double *A_val, const double *l, const double *u, const double *q) {
    if (!out) return fail(GOMP_ERR_INVALID_ARGUMENT, "out is null");
    *out = nullptr;
    if (batch <= 0 || n <= 0 || m < 0 || !P_colptr || !A_colptr) return fail(GOMP_ERR_INVALID_ARGUMENT, "bad dimensions or null pattern");
    if (batch > 0x7fffffff) return fail(GOMP_ERR_INVALID_ARGUMENT, "batch too large for one handle");
    gomp_settings st;
    if (settings) st = *settings;
    else gomp_settings_default(&st);
    if (st.check_termination < 0 || st.max_iter <= 0 || st.scaling < 0 || st.rho <= 0 || st.sigma <= 0 || st.alpha <= 0 || st.alpha >= 2)
        return fail(GOMP_ERR_INVALID_ARGUMENT, "invalid OSQP settings");
    // settings the kernels do not implement are refused instead of being ignored
    if (st.adaptive_rho && (st.adaptive_rho_interval <= 0 || (st.check_termination > 0 && st.adaptive_rho_interval % st.check_termination != 0)))
        return fail(GOMP_ERR_INVALID_ARGUMENT, "adaptive_rho_interval must be > 0 and a multiple of check_termination (OSQP's wall-clock-derived "
                                               "interval, the meaning of 0 there, is machine dependent and not reproduced)");
    if (st.adaptive_rho && st.adaptive_rho_tolerance < 1.0) return fail(GOMP_ERR_INVALID_ARGUMENT, "adaptive_rho_tolerance must be >= 1");
    if (!st.warm_start)
        return fail(GOMP_ERR_UNSUPPORTED, "warm_start = 0 is not supported: iterates persist across the solves of a handle; gomp_batch_setup is the cold start");
    if (st.threads_per_cta != 0) return fail(GOMP_ERR_UNSUPPORTED, "threads_per_cta is fixed per back end; leave it 0");
    if (gomp_device_count() <= 0) return fail(GOMP_ERR_NO_DEVICE, "no CUDA device visible: libgomp_cuda has no CPU fallback");
    if (st.linsys == GOMP_LINSYS_PCG)
        return fail(GOMP_ERR_UNSUPPORTED, "the PCG back end is not built: long horizons use the exact streamed block-tridiagonal "
                                          "factorisation (GOMP_LINSYS_BLOCKTRI_GLOBAL), which keeps OSQP's iteration counts");
    if (st.linsys < GOMP_LINSYS_AUTO || st.linsys > GOMP_LINSYS_BLOCKTRI_TMEM) return fail(GOMP_ERR_INVALID_ARGUMENT, "unknown linsys");

    gomp_batch *h = new gomp_batch();
    h->st = st;
    h->batch = batch;
    h->no_predict = std::getenv("GOMP_NO_PREDICT_ORDER") != nullptr;
    // The symbolic analysis depends on the structure only (patterns of P and A, P's values, the variable order): a planner that
    // builds a QPSolver per horizon and plans again and again (the reference's time-shrink loop, ten horizons per plan) meets the
    // same structures every time, so finished analyses are kept per process, keyed by a hash of everything they were computed from.
    {
        const int64_t nnzP = P_colptr[n], nnzA_ = A_colptr[n];
        uint64_t key = 1469598103934665603ull;
        auto mix = [&key](const void *ptr, size_t bytes) {
            const unsigned char *c = static_cast<const unsigned char *>(ptr);
            size_t k = 0;
            for (; k + 8 <= bytes; k += 8) {
                uint64_t w;
                std::memcpy(&w, c + k, 8);
                key = (key ^ w) * 1099511628211ull;
                key ^= key >> 29;
            }
            for (; k < bytes; ++k) key = (key ^ c[k]) * 1099511628211ull;
        };
        const int64_t dims[5] = {n, m, nnzP, nnzA_, (int64_t)st.block_size};
        mix(dims, sizeof dims);
        mix(P_colptr, (size_t)(n + 1) * 8);
        if (P_rowidx) mix(P_rowidx, (size_t)nnzP * 8);
        if (P_val) mix(P_val, (size_t)nnzP * 8);
        mix(A_colptr, (size_t)(n + 1) * 8);
        if (A_rowidx) mix(A_rowidx, (size_t)nnzA_ * 8);
        if (st.var_order) mix(st.var_order, (size_t)n * 8);
        static std::mutex sym_mu;
        static std::map<uint64_t, std::shared_ptr<const gomp::Symbolic>> sym_cache;
        std::shared_ptr<const gomp::Symbolic> hit;
        {
            std::lock_guard<std::mutex> lock(sym_mu);
            auto it = sym_cache.find(key);
            if (it != sym_cache.end() && it->second->n == n && it->second->m == m && it->second->nnzA == nnzA_) hit = it->second;
        }
        if (hit && !std::getenv("GOMP_NO_SYMBOLIC_CACHE")) {
            h->sym = *hit;
        } else {
            if (!gomp::analyse(h->sym, n, m, P_colptr, P_rowidx, P_val, A_colptr, A_rowidx, st.var_order, st.block_size)) {
                std::string msg = h->sym.error;
                delete h;
                return fail(GOMP_ERR_UNSUPPORTED, msg);
            }
            std::lock_guard<std::mutex> lock(sym_mu);
            if (sym_cache.size() >= 64) sym_cache.clear();
            sym_cache[key] = std::make_shared<const gomp::Symbolic>(h->sym);
        }
    }
    int rc = GOMP_OK;
    auto bail = [&](int code) {
        std::string keep = g_error;
        gomp_batch_destroy(h);
        g_error = keep;
        return code;
    };
    if (st.device >= 0) {
        if (cudaSetDevice(st.device) != cudaSuccess) return bail(fail(GOMP_ERR_CUDA, "cudaSetDevice failed"));
    }
    cudaGetDevice(&h->device);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, h->device) != cudaSuccess) return bail(fail(GOMP_ERR_CUDA, "cudaGetDeviceProperties failed"));
    h->num_sms = prop.multiProcessorCount;
    if (st.stream) {
        h->stream = (cudaStream_t)st.stream;
    } else {
        if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(fail(GOMP_ERR_CUDA, "cudaStreamCreate failed"));
        h->own_stream = true;
    }
    cudaEventCreate(&h->ev0);
    cudaEventCreate(&h->ev1);
    cudaEventCreate(&h->evc0);
    cudaEventCreate(&h->evc1);
    if ((rc = upload_pattern(h))) return bail(rc);

    // ---- streamed back end: shared-memory layout (needed before the back end is chosen: prefer_throughput)
    const gomp::Symbolic &s = h->sym;
    DevStream &x = h->dx;
    size_t o = 0;
    auto take = [&o](size_t bytes) {
        size_t at = o;
        o += (bytes + 15) & ~(size_t)15;
        return (int)at;
    };
    x.slot_doubles = s.bs * s.bs + s.bs * s.ow;
    // the grouped ring (fast path of gomp_stream.cu: compile-time block size 14, uniform 2-wide off-diagonal pattern)
    // needs grp * grp_ring stage slots per warp, the generic path kStreamRing
    const bool grouped = (s.bs == 14 || s.bs == 12) && h->ds.uniform_off && s.ow <= 2 && s.owT <= 2; // block sizes with a compiled fast path
    auto layout = [&](bool small) {
        o = 0;
        x.small = small ? 1 : 0;
        x.grp = small ? kStreamSmallGroup : kStreamGroup;
        x.grp_ring = small ? kStreamSmallGroupRing : kStreamGroupRing;
        x.oRHS = take(8 * (size_t)s.n_pad);
        // generic path: as many one-stage slots as fit beside the fixed regions (at least 3, at most kStreamRing)
        x.gen_ring = kStreamRing;
        if (!grouped) {
            const size_t fixed = 8 * (size_t)s.n_pad + 8 * ((size_t)3 * s.bs * s.bs + (size_t)4 * s.bs * std::max(1, s.ow)) + 8192;
            const size_t room = (size_t)prop.sharedMemPerBlockOptin > fixed ? (size_t)prop.sharedMemPerBlockOptin - fixed : 0;
            x.gen_ring = (int)std::max<size_t>(3, std::min<size_t>(kStreamRing, room / (2 * 8 * (size_t)x.slot_doubles)));
        }
        x.ring_doubles = (grouped ? std::max(small ? 0 : kStreamRing, x.grp * x.grp_ring) : x.gen_ring) * x.slot_doubles;
        x.oRing = take(8 * (size_t)2 * x.ring_doubles);
        const int groups = 2; // factor groups: 2 x 224 threads, or 2 x 112 in the 256-thread CTAs of the small-QP mode
        // Ruiz scratch: P values in the right-hand side region, column factors + P-side patterns in the ring region
        x.p_scratch = ((size_t)s.nnzPu <= (size_t)s.n_pad &&
                       8 * (size_t)s.n_pad + 4 * ((size_t)s.nnzPu + s.pf_csr.size()) + 2 * ((size_t)s.n_pad + 2) + 16 <= 8 * (size_t)2 * x.ring_doubles)
                          ? 1 : 0;
        x.e_scratch = 0;
        if (x.p_scratch) {
            const size_t used = (8 * (size_t)s.n_pad + 4 * ((size_t)s.nnzPu + s.pf_csr.size()) + 2 * ((size_t)s.n_pad + 2) + 15) & ~(size_t)15;
            if (used + 8 * (size_t)s.m_used <= 8 * (size_t)2 * x.ring_doubles) x.e_scratch = (int)(used / 8);
        }
        // factorisation scratch (published pivot rows, previous inverse, staged off-diagonal block).  In the small-QP
        // mode it lives inside the w = rho z - y region, which is dead while a factorisation runs (w is rebuilt after
        // every one of them)
        const size_t szPub = 8 * (size_t)128 * groups, szDprev = (8 * (size_t)groups * s.bs * s.bs + 15) & ~(size_t)15,
                     szOffBuf = (8 * (size_t)groups * s.bs * std::max(1, s.ow) + 15) & ~(size_t)15;
        const bool scratch_in_w = small && szPub + szDprev + szOffBuf <= 8 * (size_t)std::max(1, s.m_used);
        if (!scratch_in_w) {
            x.oPub = take(szPub);
            x.oDprev = take(szDprev);
            x.oOffBuf = take(szOffBuf);
        }
        x.oV = take(8 * 128);
        x.oMid = take(8 * ((size_t)s.bs * s.bs + 2 * (size_t)s.bs * std::max(1, s.ow)));
        x.oRed = take(8 * (size_t)kRedK * ((small ? kStreamSmallThreads : kStreamThreads) / 32));
        x.oBar = take(8 * 4 * kStreamRing + 16 + 64); // full + empty barriers of both chains, work-queue slot, rho table
        x.oA = x.oW = x.oRowEnt = x.oRowRange = x.oColRange = x.oCscRow = 0;
        if (small) {
            x.oA = take(8 * (size_t)s.nnzA);
            x.oW = take(8 * (size_t)std::max(1, s.m_used));
            if (scratch_in_w) {
                x.oPub = x.oW;
                x.oDprev = x.oPub + (int)szPub;
                x.oOffBuf = x.oDprev + (int)szDprev;
            }
            x.oRowEnt = take(4 * ((size_t)s.nnzA + 4));
            x.oRowRange = take(2 * ((size_t)s.m_used + 2)); // 16-bit CSR row pointers
            x.oColRange = take(2 * (size_t)s.n_pad); // 16 bits per column: first entry | length << 12
            x.oCscRow = take(2 * ((size_t)s.nnzA + 4));
        }
        return o;
    };
    // small-QP mode: two 256-thread CTAs per SM, sparse data in shared memory -- when a CTA fits half an SM
    const size_t half_sm = (size_t)prop.sharedMemPerMultiprocessor / 2 - 1024;
    bool cols_pack = true; // 16-bit packed column ranges: first entry < 4096, at most 15 entries
    for (uint32_t cr : s.col_range) cols_pack = cols_pack && (cr >> 8) < 4096u && (cr & 0xFFu) <= 15u;
    const bool small_ok = s.bs == 14 && grouped && !std::getenv("GOMP_STREAM_NO_SMALL") && layout(true) <= half_sm &&
                          cols_pack && s.m_used <= 16 * kStreamSmallThreads; // a row's rho kind: two bits of a 32-bit register per owner thread

    // ---- two-slot back end (gomp_duo.cu): eligibility and shared-memory layout
    std::string duo_why;
    size_t duo_smem = 0;
    {
        DevDuo &dd = h->dd;
        const int hmid = s.T / 2;
        gomp::build_owner_ell(s, kDuoOwners, h->oell);
        const gomp::OwnerEll &oe = h->oell;
        if (!(h->ds.uniform_off && s.ow <= 2 && s.owT <= 2)) duo_why = "off-diagonal blocks are not uniform with at most two entries per row and column";
        else if (s.T < 4 || (hmid + 1) * 16 > kDuoTmemCols) duo_why = "needs 4 <= number of stages <= 63 (one TMEM column group per stage and chain)";
        else if (oe.rpt > kDuoMaxRpt || oe.cpt > kDuoMaxCpt || !duo_variant_exists(oe.rpt, oe.cpt, s.bs)) duo_why = "no kernel variant for this block size / rows per owner";
        else if (*std::max_element(oe.plen.begin(), oe.plen.end()) > 4 || (s.nnzPu + kDuoOwners - 1) / kDuoOwners > 4)
            duo_why = "P has more than 4 entries per column or more than 4 per owner thread (register-held P patterns)";
        else if (s.nnzA > 8190 || s.m_used > 8189 || s.n_pad > 8189 || s.nnzPu > 8190) duo_why = "more than 8190 entries / rows / variables (16-bit byte offsets)";
        else {
            dd.rpt = oe.rpt; dd.cpt = oe.cpt;
            for (int j = 0; j < kDuoMaxRpt; ++j) { dd.rlen[j] = j < oe.rpt ? oe.rlen[j] : 0; dd.rbase[j] = j < oe.rpt ? oe.rbase[j] * kDuoOwners * 4 : 0; } // (byte offset of the slot's first entry line)
            for (int c = 0; c < kDuoMaxCpt; ++c) {
                dd.clen[c] = c < oe.cpt ? oe.clen[c] : 0; dd.cbase[c] = c < oe.cpt ? oe.cbase[c] * kDuoOwners * 4 : 0;
                dd.plen[c] = c < oe.cpt ? oe.plen[c] : 0; dd.pbase[c] = c < oe.cpt ? oe.pbase[c] * kDuoOwners : 0;
            }
            dd.n_row_ell = (int)oe.row_ell.size(); dd.n_col_ell = (int)oe.col_ell.size();
            size_t o2 = 0;
            auto take2 = [&o2](size_t bytes) {
                size_t at = o2;
                o2 += (bytes + 15) & ~(size_t)15;
                return (int)at;
            };
            dd.oRowEll = take2(4 * oe.row_ell.size());
            dd.oColEll = take2(4 * oe.col_ell.size());
            dd.slot_base = (int)o2;
            const size_t mu1 = (size_t)std::max(1, s.m_used);
            auto slot_layout = [&](bool dedicated) {
                o2 = 0;
                dd.oA = take2(8 * ((size_t)s.nnzA + 2));
                dd.oP = take2(8 * ((size_t)s.nnzPu + 2));
                dd.oW = take2(8 * (mu1 + 2));
                dd.oRHS = take2(8 * ((size_t)s.n_pad + 2));
                dd.oLo = take2(8 * mu1);
                dd.oHi = take2(8 * mu1);
                dd.oOff = take2(8 * (size_t)std::max(1, h->ds.noff));
                dd.oV = take2(512);
                dd.oRed = take2(8 * (size_t)kRedK * (kDuoOwners / 32));
                dd.oCtl = take2(128); // control word, work slot, mbarrier, cost scale, rho tables
                if (dedicated) {
                    dd.oScr = take2(duo_scratch_bytes(s.bs));
                    dd.oRing = take2(duo_ring_bytes(s.bs));
                } else { // the factorisation scratch and the S staging ring share the right-hand side region (dead while factoring)
                    dd.oScr = dd.oRHS;
                    dd.oRing = dd.oRHS + (int)((duo_scratch_bytes(s.bs) + 15) & ~(size_t)15);
                }
                dd.slot_bytes = (int)o2;
                return (size_t)dd.slot_base + (size_t)kDuoSlots * o2;
            };
            duo_smem = slot_layout(true);
            if (duo_smem > (size_t)prop.sharedMemPerBlockOptin) {
                const bool alias_ok = ((duo_scratch_bytes(s.bs) + 15) & ~(size_t)15) + duo_ring_bytes(s.bs) <= 8 * (size_t)s.n_pad;
                duo_smem = slot_layout(false);
                if (!alias_ok || duo_smem > (size_t)prop.sharedMemPerBlockOptin)
                    duo_why = "two QPs do not fit the shared memory of one SM (" + std::to_string(duo_smem) + " bytes needed)";
            }
        }
    }
    const bool duo_ok = duo_why.empty();
    if (st.linsys == GOMP_LINSYS_BLOCKTRI_TMEM && !duo_ok) return bail(fail(GOMP_ERR_UNSUPPORTED, "two-slot TMEM back end: " + duo_why));
    h->duo = st.linsys == GOMP_LINSYS_BLOCKTRI_TMEM || (st.linsys == GOMP_LINSYS_AUTO && duo_ok && !std::getenv("GOMP_NO_DUO"));

    h->kernel = pick_kernel(h->sym.m_used, h->sym.n_pad, h->sym.bs);
    const bool resident_ok = h->kernel.launch && h->smem_bytes <= (size_t)prop.sharedMemPerBlockOptin;
    h->streamed = !h->duo && (st.linsys == GOMP_LINSYS_BLOCKTRI_GLOBAL ||
                              (st.linsys == GOMP_LINSYS_AUTO && (!resident_ok || (st.prefer_throughput && small_ok && batch >= 2 * (int64_t)h->num_sms))));
    int occ = 0;
    if (h->duo) {
        h->smem_bytes = duo_smem;
        cudaError_t e = prepare_duo(h->dd.rpt, h->dd.cpt, s.bs, h->smem_bytes, &occ);
        if (e != cudaSuccess) return bail(fail(GOMP_ERR_CUDA, std::string("kernel preparation failed: ") + cudaGetErrorString(e)));
        if (occ < 1) return bail(fail(GOMP_ERR_CUDA, "kernel cannot be resident (occupancy 0)"));
        occ = 1; // the CTA owns the SM's whole tensor memory
        if (h->d_rell.upload(h->oell.row_ell) != cudaSuccess || h->d_cell.upload(h->oell.col_ell) != cudaSuccess ||
            h->d_pell.upload(h->oell.pf_ell) != cudaSuccess || h->d_cpos.upload(h->oell.col_pos) != cudaSuccess)
            return bail(fail(GOMP_ERR_ALLOC, "cudaMalloc failed for the owner ELL views"));
        h->dd.row_ell = h->d_rell.p; h->dd.col_ell = h->d_cell.p; h->dd.pf_ell = h->d_pell.p; h->dd.col_pos = h->d_cpos.p;
        gomp::build_assembly_ell(s, kDuoOwners, h->aell);
        {
            std::vector<unsigned long long> c1(h->aell.con.begin(), h->aell.con.end()), c2(h->aell.ocon.begin(), h->aell.ocon.end());
            if (h->d_ascon.upload(c1) != cudaSuccess || h->d_asocon.upload(c2) != cudaSuccess || h->d_asmeta.upload(h->aell.meta) != cudaSuccess ||
                h->d_asometa.upload(h->aell.ometa) != cudaSuccess)
                return bail(fail(GOMP_ERR_ALLOC, "cudaMalloc failed for the assembly lists"));
        }
        h->dd.as_con = h->d_ascon.p; h->dd.as_ocon = h->d_asocon.p; h->dd.as_meta = h->d_asmeta.p; h->dd.as_ometa = h->d_asometa.p;
        h->dd.as_ept = h->aell.ept; h->dd.as_maxc = h->aell.maxc; h->dd.as_opt = h->aell.opt; h->dd.as_maxco = h->aell.maxco;
    } else if (!h->streamed) {
        if (!h->kernel.launch)
            return bail(fail(GOMP_ERR_UNSUPPORTED, "problem too large for the register-resident kernel (non-empty rows <= 3584, variables <= 1792)"));
        if (h->smem_bytes > (size_t)prop.sharedMemPerBlockOptin)
            return bail(fail(GOMP_ERR_UNSUPPORTED, "QP does not fit in shared memory (" + std::to_string(h->smem_bytes) + " bytes needed)"));
        cudaError_t e = h->kernel.prepare(h->smem_bytes, &occ);
        if (e != cudaSuccess) return bail(fail(GOMP_ERR_CUDA, std::string("kernel preparation failed: ") + cudaGetErrorString(e)));
        if (occ < 1) return bail(fail(GOMP_ERR_CUDA, "kernel cannot be resident (occupancy 0)"));
    } else {
        // long-horizon back end: x~ in shared memory, everything else in a per-CTA global workspace
        if (s.bs % 2 != 0 || s.bs > 32)
            return bail(fail(GOMP_ERR_UNSUPPORTED, "streamed back end needs an even block size <= 32 (16-byte bulk copies)"));
        const bool small = small_ok;
        h->smem_bytes = layout(small);
        if (h->smem_bytes > (size_t)prop.sharedMemPerBlockOptin)
            return bail(fail(GOMP_ERR_UNSUPPORTED, "right-hand side of the block solve does not fit in shared memory (" +
                                                       std::to_string(h->smem_bytes) + " bytes needed)"));
        cudaError_t e = prepare_stream(s.bs, small, h->smem_bytes, &occ);
        if (e != cudaSuccess) return bail(fail(GOMP_ERR_CUDA, std::string("kernel preparation failed: ") + cudaGetErrorString(e)));
        if (occ < 1) return bail(fail(GOMP_ERR_CUDA, "kernel cannot be resident (occupancy 0)"));
        long long w = 0;
        auto wtake = [&w](long long cnt) {
            long long at = w;
            w += (cnt + 1) & ~1LL;
            return at;
        };
        const long long mu1 = std::max(1, s.m_used);
        x.wA = wtake(s.nnzA); x.wP = wtake(std::max<long long>(1, s.nnzPu)); x.wW = wtake(mu1); x.wRho = wtake(mu1);
        x.wDinv = wtake((long long)s.T * s.bs * s.bs); x.wOff = wtake(std::max<long long>(1, (long long)std::max(0, s.T - 1) * s.bs * s.ow));
        x.wT = wtake(s.n_pad); x.wLo = wtake(mu1); x.wHi = wtake(mu1);
        x.stride = w;
        if (h->d_row_ell.upload(s.row_ent) != cudaSuccess || h->d_col_ell.upload(s.col_row) != cudaSuccess)
            return bail(fail(GOMP_ERR_ALLOC, "cudaMalloc failed for the ELL pattern views"));
        x.row_ell = h->d_row_ell.p; x.col_ell = h->d_col_ell.p; x.roww = s.roww; x.colw = s.colw;
        for (int k = 0; k < 256; ++k) x.rlen_blk[k] = 0;
        for (int i = 0; i < s.m_used; ++i) {
            const int k = std::min(i >> 8, 255);
            x.rlen_blk[k] = (unsigned char)std::max<int>(x.rlen_blk[k], (int)(s.row_range[(size_t)i] & 0xFFu));
        }
        for (int k = 254; k >= 0; --k) x.rlen_blk[k] = std::max(x.rlen_blk[k], x.rlen_blk[k + 1]); // (monotone even if the order were not)
    }
    h->ctas_per_sm = occ;
    h->grid = (int)std::min<int64_t>(batch, (int64_t)h->num_sms * occ);

    // per-batch storage
    const size_t B = (size_t)batch, np = (size_t)h->sym.n_pad, mu = (size_t)std::max(1, h->sym.m_used);
#define ALLOC(arr, cnt)                                                                            \
    do {                                                                                           \
        if (h->arr.alloc(cnt) != cudaSuccess) return bail(fail(GOMP_ERR_ALLOC, "cudaMalloc failed for " #arr)); \
        h->workspace_bytes += (size_t)(cnt) * sizeof(*h->arr.p);                                   \
    } while (0)
    ALLOC(d_x0, B * (size_t)n);
    ALLOC(d_xs, B * np);
    ALLOC(d_zs, B * mu);
    ALLOC(d_ys, B * mu);
    ALLOC(d_rho, B);
    ALLOC(d_E, B * mu);
    ALLOC(d_D, B * np);
    ALLOC(d_dx, B * np);
    if (h->streamed) {
        // Latency mode: when the batch leaves SMs idle, a QP is served by a CLUSTER of CTAs -- 8 per QP up to SMs / 8 QPs, then 4,
        // then 2: the Ruiz equilibration runs as a cluster kernel ahead of the ADMM kernel (gomp_ruiz_cluster_kernel), and the ADMM
        // kernel itself is launched with that cluster size, the extra CTAs sharing the sparse passes of every iteration
        // (gomp_stream.cu).  Larger batches keep every SM busy with its own QP.  GOMP_RUIZ_CLUSTER=<0|1|2|4|8> overrides (A/B runs).
        h->dx.pre = nullptr;
        h->pre_cluster = 0;
        if (!h->dx.small && h->st.scaling > 0) {
            int cs = 0;
            for (int c : {8, 4, 2}) {
                int fit = 0; // clusters of c CTAs the device holds at once (a cluster lives inside one GPC: fewer than SMs / c)
                if (cs == 0 && stream_max_clusters(s.bs, c, h->smem_bytes, &fit) == cudaSuccess && batch <= (int64_t)fit) {
                    cs = c;
                    h->max_clusters = fit;
                }
                cudaGetLastError();
            }
            if (const char *e = std::getenv("GOMP_RUIZ_CLUSTER")) {
                const int v = std::atoi(e);
                cs = (v == 1 || v == 2 || v == 4 || v == 8) ? v : 0;
                int fit = 0; // a forced size: as many clusters as are co-resident, the leaders then take several QPs each
                h->max_clusters = (cs > 0 && stream_max_clusters(s.bs, cs, h->smem_bytes, &fit) == cudaSuccess) ? fit : 0;
                cudaGetLastError();
            }
            h->pre_cluster = cs;
        }
        h->solve_cluster = h->pre_cluster > 1 && !std::getenv("GOMP_NO_SOLVE_CLUSTER") ? h->pre_cluster : 1;
        if (h->solve_cluster > 1) h->grid = (int)std::max<int64_t>(1, std::min<int64_t>(batch, h->max_clusters > 0 ? h->max_clusters : h->num_sms / h->solve_cluster)) * h->solve_cluster;
        ALLOC(d_ws, (size_t)h->grid * (size_t)h->dx.stride);
        h->dx.ws = h->d_ws.p;
        if (h->pre_cluster > 0) {
            DevStream &x = h->dx;
            long long w = 0;
            auto wtake = [&w](long long cnt) {
                long long at = w;
                w += (cnt + 1) & ~1LL;
                return at;
            };
            x.pA = wtake(s.nnzA); x.pP = wtake(std::max<long long>(1, s.nnzPu)); x.pLo = wtake((long long)mu); x.pHi = wtake((long long)mu);
            x.pT = wtake((long long)np); x.pW = wtake((long long)mu); x.pRed = wtake(32); x.pC = wtake(2);
            x.pre_stride = w;
            ALLOC(d_pre, B * (size_t)w);
            x.pre = h->d_pre.p;
        }
    } else if (!h->ds.bounds_in_smem) {
        ALLOC(d_Lg, B * mu);
        ALLOC(d_Ug, B * mu);
    }
    ALLOC(d_dy, B * mu);
    ALLOC(d_xo, B * (size_t)n);
    ALLOC(d_yo, B * (size_t)std::max<int64_t>(m, 1));
    ALLOC(d_pri, B);
    ALLOC(d_dua, B);
    ALLOC(d_obj, B);
    ALLOC(d_status, B);
    ALLOC(d_iters, B);
    ALLOC(d_rup, B);
    ALLOC(d_counter, 1);
    if (std::getenv("GOMP_PROFILE_PHASES")) {
        ALLOC(d_prof, 16);
        cudaMemsetAsync(h->d_prof.p, 0, 16 * sizeof(unsigned long long), h->stream);
    }
#undef ALLOC
    cudaMemsetAsync(h->d_rup.p, 0, B * sizeof(int), h->stream);
    cudaMemsetAsync(h->d_iters.p, 0, B * sizeof(int), h->stream);
    cudaMemsetAsync(h->d_status.p, 0, B * sizeof(int), h->stream);

    h->dst.rho = st.rho; h->dst.sigma = st.sigma; h->dst.alpha = st.alpha;
    h->dst.adaptive_rho_tolerance = st.adaptive_rho_tolerance;
    h->dst.eps_abs = st.eps_abs; h->dst.eps_rel = st.eps_rel; h->dst.eps_prim_inf = st.eps_prim_inf; h->dst.eps_dual_inf = st.eps_dual_inf;
    h->dst.scaling = st.scaling; h->dst.adaptive_rho = st.adaptive_rho; h->dst.adaptive_rho_interval = st.adaptive_rho_interval;
    h->dst.max_iter = st.max_iter; h->dst.scaled_termination = st.scaled_termination; h->dst.check_termination = st.check_termination;

    if (A_val || l || u) {
        if ((rc = load_data(h, A_val, l, u, q, true))) return bail(rc);
    }
    *out = h;
    return GOMP_OK;
}

int gomp_batch_setup(gomp_batch *h, const double *A_val, const double *l, const double *u, const double *q) {
    if (!h) return fail(GOMP_ERR_INVALID_ARGUMENT, "null handle");
    return load_data(h, A_val, l, u, q, true);
}

int gomp_batch_update(gomp_batch *h, const double *A_val, const double *l, const double *u, const double *q) {
    if (!h) return fail(GOMP_ERR_INVALID_ARGUMENT, "null handle");
    if (!h->data_loaded) return fail(GOMP_ERR_INVALID_ARGUMENT, "update before setup");
    if ((l == nullptr) != (u == nullptr)) return fail(GOMP_ERR_INVALID_ARGUMENT, "l and u must be updated together");
    return load_data(h, A_val, l, u, q, false);
}

int gomp_batch_warm_start_x(gomp_batch *h, const double *x0) {
    if (!h || !x0) return fail(GOMP_ERR_INVALID_ARGUMENT, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t B = (size_t)h->batch;
    if (is_device_pointer(x0) && h->st.borrow_inputs) {
        h->x0_src = x0;
    } else {
        CUDA_TRY(cudaMemcpyAsync(h->d_x0.p, x0, B * (size_t)h->sym.n * sizeof(double),
                                 is_device_pointer(x0) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
        h->x0_src = h->d_x0.p;
    }
    h->pending_flags |= 2;
    return GOMP_OK;
}

int gomp_batch_solve(gomp_batch *h) {
    if (!h) return fail(GOMP_ERR_INVALID_ARGUMENT, "null handle");
    if (!h->data_loaded) return fail(GOMP_ERR_INVALID_ARGUMENT, "solve before setup");
    CUDA_TRY(cudaSetDevice(h->device));
    const gomp::Symbolic &s = h->sym;
    DevBatch &b = h->db;
    b.batch = (int)h->batch;
    b.A_val = h->A_src; b.l = h->l_src; b.u = h->u_src; b.q = h->has_q ? h->q_src : nullptr; b.x0 = h->x0_src;
    b.A_pitch = s.nnzA;
    b.bulk_ok = (s.nnzA % 2 == 0) && (((uintptr_t)h->A_src) % 16 == 0) && s.nnzA > 0 && (size_t)s.nnzA * 8 < (1u << 20);
    b.x_state = h->d_xs.p; b.z_state = h->d_zs.p; b.y_state = h->d_ys.p; b.rho_state = h->d_rho.p; b.flag_bits = h->pending_flags;
    b.Eg = h->d_E.p; b.Dg = h->d_D.p; b.qs = h->d_qs.p; b.dyg = h->d_dy.p; b.dxg = h->d_dx.p; b.Lg = h->d_Lg.p; b.Ug = h->d_Ug.p; b.x_out = h->d_xo.p; b.y_out = h->d_yo.p;
    b.status = h->d_status.p; b.iters = h->d_iters.p; b.rho_updates = h->d_rup.p;
    b.pri_res = h->d_pri.p; b.dua_res = h->d_dua.p; b.obj_val = h->d_obj.p; b.work_counter = h->d_counter.p; b.prof = h->d_prof.p; b.active = h->active_src;
    // cold solve of new problems from a primal warm start: queue order from the warm start's constraint violation (see
    // predict_key_kernel) when the batch holds more QPs than are resident at once; GOMP_NO_PREDICT_ORDER=1 keeps index order.
    // Warm re-solve of the problems of the previous solve (gomp_batch_update): queue order from their iteration counts.
    const bool predict = !h->order_valid && (h->pending_flags & 3) == 3 && h->x0_src != nullptr &&
                         h->batch > (int64_t)2 * h->num_sms * std::max(1, h->ctas_per_sm) && !h->no_predict;
    const bool reorder = !h->order_valid && !predict && h->iters_fresh;
    if (predict || reorder) {
        if (predict && !h->d_pred.p) CUDA_TRY(h->d_pred.alloc((size_t)h->batch));
        if (!h->d_order.p) CUDA_TRY(h->d_order.alloc((size_t)h->batch));
        cudaStream_t ks = h->stream;
        if (predict) {
            const unsigned blocks = (unsigned)((h->batch * 32 + 255) / 256);
            predict_key_kernel<<<blocks, 256, 0, ks>>>(h->ds, h->A_src, (long long)s.nnzA, h->l_src, h->u_src, h->x0_src, (int)h->batch, h->d_pred.p);
            h->launches += 1;
        }
        order_by_iters_kernel<<<1, 1024, 0, ks>>>(predict ? h->d_pred.p : h->d_iters.p, (int)h->batch, h->d_order.p);
        CUDA_TRY(cudaGetLastError());
        h->launches += 1;
        h->order_valid = true;
    }
    b.order = h->order_valid ? h->d_order.p : nullptr;
    CUDA_TRY(cudaMemsetAsync(h->d_counter.p, 0, sizeof(int), h->stream));
    CUDA_TRY(cudaEventRecord(h->ev0, h->stream));
    const int grid = h->max_ctas > 0 && h->solve_cluster <= 1 ? std::min(h->grid, h->max_ctas) : h->grid;
    if (h->duo) CUDA_TRY(launch_duo(h->dd.rpt, h->dd.cpt, s.bs, h->ds, h->db, h->dst, h->dd, grid, h->smem_bytes, h->stream));
    else if (h->streamed) {
        if (h->pre_cluster > 0) {
            const int clusters = (int)std::max<int64_t>(1, std::min<int64_t>(h->batch, h->max_clusters > 0 ? h->max_clusters : h->num_sms / h->pre_cluster));
            CUDA_TRY(launch_prescale(h->ds, h->db, h->dst, h->dx, clusters, h->pre_cluster, h->stream));
            h->launches += 1;
        }
        CUDA_TRY(launch_stream(h->ds, h->db, h->dst, h->dx, grid, h->solve_cluster, h->smem_bytes, h->stream));
    }
    else CUDA_TRY(h->kernel.launch(h->ds, h->db, h->dst, grid, h->smem_bytes, h->stream));
    CUDA_TRY(cudaEventRecord(h->ev1, h->stream));
    // (the queue order of a warm re-solve of these problems and the exit codes are derived when they are asked for: a kernel
    //  queued here would delay the batch's results until an SM is free for it)
    h->order_valid = false;
    h->iters_fresh = true;
    h->code_valid = false;
    h->launches += 1;
    h->pending_flags = 0;
    return GOMP_OK;
}

int gomp_batch_set_max_ctas(gomp_batch *h, int32_t max_ctas) {
    if (!h) return fail(GOMP_ERR_INVALID_ARGUMENT, "null handle");
    if (max_ctas < 0) return fail(GOMP_ERR_INVALID_ARGUMENT, "max_ctas must be >= 0");
    h->max_ctas = max_ctas;
    return GOMP_OK;
}

int gomp_batch_set_active(gomp_batch *h, const int32_t *active) {
    if (!h) return fail(GOMP_ERR_INVALID_ARGUMENT, "null handle");
    CUDA_TRY(cudaSetDevice(h->device));
    if (!active) {
        h->active_src = nullptr;
        return GOMP_OK;
    }
    if (is_device_pointer(active) && h->st.borrow_inputs) {
        h->active_src = active;
        return GOMP_OK;
    }
    if (!h->d_active.p) CUDA_TRY(h->d_active.alloc((size_t)h->batch));
    CUDA_TRY(cudaMemcpyAsync(h->d_active.p, active, (size_t)h->batch * sizeof(int),
                             is_device_pointer(active) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->stream));
    h->active_src = h->d_active.p;
    return GOMP_OK;
}

int gomp_batch_device_views(gomp_batch *h, double **A_val, double **l, double **u, const double **x, void **stream) {
    if (!h) return fail(GOMP_ERR_INVALID_ARGUMENT, "null handle");
    if (A_val) *A_val = h->A_src == h->d_A.p ? h->d_A.p : nullptr;
    if (l) *l = h->l_src == h->d_l.p ? h->d_l.p : nullptr;
    if (u) *u = h->u_src == h->d_u.p ? h->d_u.p : nullptr;
    if (x) *x = h->d_xo.p;
    if (stream) *stream = (void *)h->stream;
    return GOMP_OK;
}

int gomp_device_malloc(void **ptr, int64_t bytes, int32_t device) {
    if (!ptr || bytes < 0) return fail(GOMP_ERR_INVALID_ARGUMENT, "bad arguments");
    if (gomp_device_count() <= 0) return fail(GOMP_ERR_NO_DEVICE, "no CUDA device visible");
    if (device >= 0) CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaMalloc(ptr, (size_t)std::max<int64_t>(bytes, 1)));
    return GOMP_OK;
}

int gomp_device_free(void *ptr) {
    if (ptr) CUDA_TRY(cudaFree(ptr));
    return GOMP_OK;
}

int gomp_device_memcpy(void *dst, const void *src, int64_t bytes, void *stream) {
    if (!dst || !src || bytes < 0) return fail(GOMP_ERR_INVALID_ARGUMENT, "bad arguments");
    CUDA_TRY(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    return GOMP_OK;
}

// wait for the handle's stream; report device-resident bounds that violated l <= u (load_data's device-side check)
static int wait_and_check(gomp_batch *h) {
    if (h->d_flag.p) CUDA_TRY(cudaMemcpyAsync(&h->h_flag, h->d_flag.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (h->d_flag.p && h->h_flag)
        return fail(GOMP_ERR_INVALID_ARGUMENT, "lower bound exceeds upper bound in the device-resident l / u of the last setup / update "
                                               "(checked on the device; the results of solves on that data are meaningless)");
    return GOMP_OK;
}

int gomp_batch_sync(gomp_batch *h) {
    if (!h) return fail(GOMP_ERR_INVALID_ARGUMENT, "null handle");
    CUDA_TRY(cudaSetDevice(h->device));
    return wait_and_check(h);
}

static int copy_out(gomp_batch *h, void *dst, const void *src, size_t bytes) {
    if (!dst) return GOMP_OK;
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, is_device_pointer(dst) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
    return GOMP_OK;
}

// `host_map`: the exit codes go to a host array and the caller waits for the stream: copy OSQP's status values and map them
// on the host afterwards (gomp_batch_get) -- no kernel, hence no wait for an SM while other batches' CTAs hold them all
static int queue_results(gomp_batch *h, double *x, double *y, int32_t *exit_code, int32_t *iters, int32_t *rho_updates, bool host_map) {
    const size_t B = (size_t)h->batch;
    int rc;
    if (exit_code && h->launches == 0) return fail(GOMP_ERR_INVALID_ARGUMENT, "no solve has been run on this handle");
    if ((rc = copy_out(h, x, h->d_xo.p, B * (size_t)h->sym.n * sizeof(double)))) return rc;
    if ((rc = copy_out(h, y, h->d_yo.p, B * (size_t)h->sym.m * sizeof(double)))) return rc;
    if ((rc = copy_out(h, iters, h->d_iters.p, B * sizeof(int)))) return rc;
    if ((rc = copy_out(h, rho_updates, h->d_rup.p, B * sizeof(int)))) return rc;
    if (exit_code && host_map) return copy_out(h, exit_code, h->d_status.p, B * sizeof(int));
    if (exit_code) {
        if (!h->code_valid) {
            if (!h->d_code.p) CUDA_TRY(h->d_code.alloc(B));
            exit_code_kernel<<<(unsigned)std::min<size_t>((B + 255) / 256, 64), 256, 0, h->stream>>>(h->d_status.p, (int)B, h->d_code.p);
            CUDA_TRY(cudaGetLastError());
            h->launches += 1;
            h->code_valid = true;
        }
        if ((rc = copy_out(h, exit_code, h->d_code.p, B * sizeof(int)))) return rc;
    }
    return GOMP_OK;
}

int gomp_batch_get_async(gomp_batch *h, double *x, double *y, int32_t *exit_code, int32_t *iters, int32_t *rho_updates) {
    if (!h) return fail(GOMP_ERR_INVALID_ARGUMENT, "null handle");
    CUDA_TRY(cudaSetDevice(h->device));
    return queue_results(h, x, y, exit_code, iters, rho_updates, false);
}

int gomp_batch_get(gomp_batch *h, double *x, double *y, int32_t *exit_code, int32_t *iters, int32_t *rho_updates) {
    if (!h) return fail(GOMP_ERR_INVALID_ARGUMENT, "null handle");
    CUDA_TRY(cudaSetDevice(h->device));
    int rc;
    const bool host_map = exit_code != nullptr && !is_device_pointer(exit_code);
    if ((rc = queue_results(h, x, y, exit_code, iters, rho_updates, host_map))) return rc;
    if ((rc = wait_and_check(h))) return rc;
    if (host_map)
        for (int64_t b = 0; b < h->batch; ++b) exit_code[b] = exit_code_of(exit_code[b]);
    float ms = 0;
    if (h->launches > 0 && cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) h->last_solve_ms = ms;
    else cudaGetLastError();
    return GOMP_OK;
}

int gomp_batch_get_info(gomp_batch *h, double *pri_res, double *dua_res, double *rho, double *obj_val) {
    if (!h) return fail(GOMP_ERR_INVALID_ARGUMENT, "null handle");
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t B = (size_t)h->batch;
    int rc;
    if ((rc = copy_out(h, pri_res, h->d_pri.p, B * sizeof(double)))) return rc;
    if ((rc = copy_out(h, dua_res, h->d_dua.p, B * sizeof(double)))) return rc;
    if ((rc = copy_out(h, rho, h->d_rho.p, B * sizeof(double)))) return rc;
    if ((rc = copy_out(h, obj_val, h->d_obj.p, B * sizeof(double)))) return rc;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return GOMP_OK;
}

int gomp_batch_get_certificates(gomp_batch *h, double *prim_inf_cert, double *dual_inf_cert) {
    if (!h) return fail(GOMP_ERR_INVALID_ARGUMENT, "null handle");
    if (h->launches == 0) return fail(GOMP_ERR_INVALID_ARGUMENT, "no solve has been run on this handle");
    if (!prim_inf_cert && !dual_inf_cert) return GOMP_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t B = (size_t)h->batch;
    const gomp::Symbolic &s = h->sym;
    // device outputs are written in place; host outputs go through a staging buffer on the handle's device
    double *dp = nullptr, *dd = nullptr;
    DevArray<double> tp, td;
    if (prim_inf_cert) {
        if (is_device_pointer(prim_inf_cert)) dp = prim_inf_cert;
        else {
            CUDA_TRY(tp.alloc(B * (size_t)s.m));
            dp = tp.p;
        }
    }
    if (dual_inf_cert) {
        if (is_device_pointer(dual_inf_cert)) dd = dual_inf_cert;
        else {
            CUDA_TRY(td.alloc(B * (size_t)s.n));
            dd = td.p;
        }
    }
    certificate_kernel<<<(unsigned)B, 256, 0, h->stream>>>((int)s.m, (int)s.n, s.m_used, s.n_pad, h->d_used_rows.p, h->d_pos_var.p, h->d_status.p,
                                                          h->d_iters.p, h->d_empty_rows.p, (int)s.empty_rows.size(), h->l_src, h->u_src, h->d_E.p, h->d_D.p, h->d_dy.p, h->has_q ? h->d_dx.p : nullptr,
                                                          h->st.scaled_termination, dp, dd);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && tp.p) e = cudaMemcpyAsync(prim_inf_cert, tp.p, B * (size_t)s.m * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess && td.p) e = cudaMemcpyAsync(dual_inf_cert, td.p, B * (size_t)s.n * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    tp.release();
    td.release();
    if (e != cudaSuccess) return fail(GOMP_ERR_CUDA, std::string("certificates: ") + cudaGetErrorString(e));
    h->launches += 1;
    return GOMP_OK;
}

int gomp_batch_info(gomp_batch *h, gomp_info *info) {
    if (!h || !info) return fail(GOMP_ERR_INVALID_ARGUMENT, "null argument");
    std::memset(info, 0, sizeof *info);
    const gomp::Symbolic &s = h->sym;
    info->batch = h->batch; info->n = s.n; info->m = s.m; info->m_used = s.m_used; info->nnzA = s.nnzA; info->nnzP = s.nnzPu;
    info->block_size = s.bs; info->num_blocks = s.T; info->offdiag_width = s.ow;
    info->linsys = h->duo ? GOMP_LINSYS_BLOCKTRI_TMEM : (h->streamed ? GOMP_LINSYS_BLOCKTRI_GLOBAL : GOMP_LINSYS_BLOCKTRI_SMEM);
    info->threads_per_cta = h->duo ? kDuoThreads : (h->streamed ? (h->dx.small ? kStreamSmallThreads : kStreamThreads) : kThreads);
    info->ctas_per_sm = h->ctas_per_sm;
    info->qps_per_sm = h->duo ? kDuoSlots : h->ctas_per_sm;
    info->num_sms = h->num_sms; info->grid = h->grid; info->smem_bytes = (int64_t)h->smem_bytes;
    info->workspace_bytes = (int64_t)h->workspace_bytes; info->kernel_launches = h->launches;
    float ms = 0;
    if (h->launches > 0 && cudaEventQuery(h->ev1) == cudaSuccess && cudaEventElapsedTime(&ms, h->ev0, h->ev1) == cudaSuccess) h->last_solve_ms = ms;
    else cudaGetLastError();
    if (h->data_loaded && cudaEventQuery(h->evc1) == cudaSuccess && cudaEventElapsedTime(&ms, h->evc0, h->evc1) == cudaSuccess) h->last_h2d_ms = ms;
    else cudaGetLastError();
    if (h->d_prof.p) {
        unsigned long long pc[16];
        if (cudaMemcpy(pc, h->d_prof.p, sizeof pc, cudaMemcpyDeviceToHost) == cudaSuccess)
            for (int k = 0; k < 8; ++k) info->phase_cycles[k] = (double)pc[k];
            info->phase_cycles[2] += (double)pc[8]; // assembly is reported inside 'factor'
            info->assemble_cycles = (double)pc[8];
            info->chain_cycles = (double)pc[9];
            if (std::getenv("GOMP_PROFILE_PHASES")[0] == '2') {
                std::fprintf(stderr, "chain marks (cycles): [10] %.0f  [11] %.0f  [12] %.0f  total %.0f\n", (double)pc[10], (double)pc[11], (double)pc[12], (double)pc[9]);
                std::fprintf(stderr, "info marks (cycles): rows %.0f  cols %.0f  termination %.0f\n", (double)pc[13], (double)pc[14], (double)pc[15]);
            }
    }
    info->last_solve_ms = h->last_solve_ms;
    info->last_h2d_ms = h->last_h2d_ms;
    // algorithmic work of one QP (DESIGN.md "work model"); 1 FMA = 2 flops, shared-memory bytes = operands read + written
    const int64_t bs = s.bs, T = s.T, dpb = s.dpb, noff = (int64_t)std::max(0, s.T - 1) * s.bs * s.ow;
    const int64_t offnz = (int64_t)s.so_ptr.size() > 0 ? noff : 0;
    const int64_t solve_fma = 2 * (T * bs * bs + 2 * offnz); // forward + backward: one dense GEMV and one sparse product per block
    info->flops_per_iter = 2 * (solve_fma + 2 * s.nnzA) + 10 * (int64_t)s.m_used + 5 * s.n;
    info->smem_bytes_per_iter = 8 * (2 * (T * dpb + noff)      /* factor read twice */
                                     + 2 * 2 * s.nnzA          /* A value + gathered operand, two SpMVs */
                                     + 4 * s.n_pad + 2 * (int64_t)s.m_used);
    info->flops_per_factor = 2 * ((int64_t)s.sd_con.size() + (int64_t)s.so_con.size()) * 2 + 2 * T * bs * bs * bs + 2 * T * dpb * 4;
    info->smem_bytes_per_factor = 8 * (3 * ((int64_t)s.sd_con.size() + (int64_t)s.so_con.size()) + 3 * T * bs * bs * bs);
    info->flops_per_scaling = (int64_t)std::max(0, h->st.scaling) * 6 * (s.nnzA + s.nnzPu);
    info->smem_bytes_per_scaling = (int64_t)std::max(0, h->st.scaling) * 8 * 4 * (s.nnzA + s.nnzPu);
    info->flops_per_check = 2 * (2 * s.nnzA + 2 * s.nnzPu) + 20 * ((int64_t)s.m_used + s.n);
    info->smem_bytes_per_check = 8 * (2 * 2 * s.nnzA + 2 * 2 * s.nnzPu + 2 * (int64_t)s.m_used + 2 * s.n);
    return GOMP_OK;
}

} // extern "C"

// ================================================================================================
//                          one batch over several GPUs of one process
// ================================================================================================
struct MgShard {
    int device = 0;
    int64_t count = 0;
    gomp_batch *h = nullptr;
    cudaStream_t stream = nullptr; // the shard handle's stream (owned by the handle)
    double *A = nullptr, *l = nullptr, *u = nullptr, *q = nullptr, *x0 = nullptr; // staged inputs [count, .]
    double *xo = nullptr, *yo = nullptr;                                         // staged outputs
    int32_t *ints = nullptr;                                                     // code, iters, rho_updates [3][count]
    std::vector<int32_t> h_ints;
};
struct gomp_mg {
    std::vector<MgShard> sh;
    int64_t batch = 0, n = 0, m = 0, nnzA = 0;
    bool has_q = false;
};

namespace {
// rows k, k + G, k + 2G, ... of a [batch, width] array <-> a contiguous [count, width] shard buffer
int mg_rows_in(const gomp_mg *g, const MgShard &s, int k, double *dst, const double *src, int64_t width) {
    const size_t G = g->sh.size(), rb = (size_t)width * sizeof(double);
    CUDA_TRY(cudaMemcpy2DAsync(dst, rb, src + (size_t)k * width, G * rb, rb, (size_t)s.count, cudaMemcpyDefault, s.stream));
    return GOMP_OK;
}
int mg_rows_out(const gomp_mg *g, const MgShard &s, int k, double *dst, const double *src, int64_t width) {
    const size_t G = g->sh.size(), rb = (size_t)width * sizeof(double);
    CUDA_TRY(cudaMemcpy2DAsync(dst + (size_t)k * width, G * rb, src, rb, rb, (size_t)s.count, cudaMemcpyDefault, s.stream));
    return GOMP_OK;
}
int mg_stage(gomp_mg *g, MgShard &s, int k, double **buf, const double *src, int64_t width) {
    if (!src) return GOMP_OK;
    if (!*buf) CUDA_TRY(cudaMalloc((void **)buf, (size_t)std::max<int64_t>(1, s.count * width) * sizeof(double)));
    return mg_rows_in(g, s, k, *buf, src, width);
}
int mg_load(gomp_mg *g, const double *A_val, const double *l, const double *u, const double *q, bool cold) {
    if (!g) return fail(GOMP_ERR_INVALID_ARGUMENT, "null handle");
    if (cold && (!A_val || !l || !u)) return fail(GOMP_ERR_INVALID_ARGUMENT, "setup needs A_val, l and u");
    if ((l == nullptr) != (u == nullptr)) return fail(GOMP_ERR_INVALID_ARGUMENT, "l and u must be given together");
    if (l && u && !is_device_pointer(l) && !is_device_pointer(u) && !bounds_ordered(l, u, (size_t)g->batch * (size_t)g->m))
        return fail(GOMP_ERR_INVALID_ARGUMENT, "lower bound exceeds upper bound");
    for (size_t k = 0; k < g->sh.size(); ++k) {
        MgShard &s = g->sh[k];
        if (s.count == 0) continue;
        CUDA_TRY(cudaSetDevice(s.device));
        int rc;
        if ((rc = mg_stage(g, s, (int)k, &s.A, A_val, g->nnzA))) return rc;
        if ((rc = mg_stage(g, s, (int)k, &s.l, l, g->m))) return rc;
        if ((rc = mg_stage(g, s, (int)k, &s.u, u, g->m))) return rc;
        if ((rc = mg_stage(g, s, (int)k, &s.q, q, g->n))) return rc;
        // the shard handle borrows the staged arrays (settings.borrow_inputs = 1): no second copy
        rc = cold ? gomp_batch_setup(s.h, s.A, s.l, s.u, q ? s.q : nullptr)
                  : gomp_batch_update(s.h, A_val ? s.A : nullptr, l ? s.l : nullptr, u ? s.u : nullptr, q ? s.q : nullptr);
        if (rc) return rc;
    }
    return GOMP_OK;
}
} // namespace

extern "C" {

int gomp_mg_destroy(gomp_mg *g) {
    if (!g) return GOMP_OK;
    for (MgShard &s : g->sh) {
        cudaSetDevice(s.device);
        if (s.h) gomp_batch_destroy(s.h);
        for (double *p : {s.A, s.l, s.u, s.q, s.x0, s.xo, s.yo})
            if (p) cudaFree(p);
        if (s.ints) cudaFree(s.ints);
    }
    delete g;
    return GOMP_OK;
}

int gomp_mg_create(gomp_mg **out, const gomp_settings *settings, const int32_t *devices, int32_t n_devices, int64_t batch, int64_t n,
                   int64_t m, const int64_t *P_colptr, const int64_t *P_rowidx, const double *P_val, const int64_t *A_colptr,
                   const int64_t *A_rowidx) {
    if (!out) return fail(GOMP_ERR_INVALID_ARGUMENT, "out is null");
    *out = nullptr;
    if (batch <= 0 || n <= 0 || !A_colptr) return fail(GOMP_ERR_INVALID_ARGUMENT, "bad dimensions or null pattern");
    const int visible = gomp_device_count();
    if (visible <= 0) return fail(GOMP_ERR_NO_DEVICE, "no CUDA device visible: libgomp_cuda has no CPU fallback");
    if (n_devices <= 0) {
        n_devices = visible;
        devices = nullptr;
    }
    gomp_mg *g = new gomp_mg();
    g->batch = batch; g->n = n; g->m = m; g->nnzA = A_colptr[n];
    g->sh.resize((size_t)n_devices);
    gomp_settings st;
    if (settings) st = *settings;
    else gomp_settings_default(&st);
    st.stream = nullptr;
    st.borrow_inputs = 1;
    for (int k = 0; k < n_devices; ++k) {
        MgShard &s = g->sh[(size_t)k];
        s.device = devices ? devices[k] : k;
        if (s.device < 0 || s.device >= visible) {
            gomp_mg_destroy(g);
            return fail(GOMP_ERR_INVALID_ARGUMENT, "device ordinal out of range");
        }
        s.count = batch > k ? (batch - k + n_devices - 1) / n_devices : 0;
        if (s.count == 0) continue;
        st.device = s.device;
        int rc = gomp_batch_create(&s.h, &st, s.count, n, m, P_colptr, P_rowidx, P_val, A_colptr, A_rowidx, nullptr, nullptr, nullptr, nullptr);
        if (rc) {
            std::string keep = g_error;
            gomp_mg_destroy(g);
            return fail(rc, keep);
        }
        void *stream = nullptr;
        gomp_batch_device_views(s.h, nullptr, nullptr, nullptr, nullptr, &stream);
        s.stream = (cudaStream_t)stream;
    }
    // peer access between the devices of the handle (device-pointer inputs / outputs living on another GPU); not fatal when absent:
    // cudaMemcpyDefault then stages through the host
    for (const MgShard &a : g->sh)
        for (const MgShard &b : g->sh)
            if (a.device != b.device) {
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, a.device, b.device) == cudaSuccess && can) {
                    cudaSetDevice(a.device);
                    cudaDeviceEnablePeerAccess(b.device, 0);
                }
                cudaGetLastError();
            }
    *out = g;
    return GOMP_OK;
}

int gomp_mg_setup(gomp_mg *g, const double *A_val, const double *l, const double *u, const double *q) { return mg_load(g, A_val, l, u, q, true); }

int gomp_mg_update(gomp_mg *g, const double *A_val, const double *l, const double *u, const double *q) { return mg_load(g, A_val, l, u, q, false); }

int gomp_mg_warm_start_x(gomp_mg *g, const double *x0) {
    if (!g || !x0) return fail(GOMP_ERR_INVALID_ARGUMENT, "null argument");
    for (size_t k = 0; k < g->sh.size(); ++k) {
        MgShard &s = g->sh[k];
        if (s.count == 0) continue;
        CUDA_TRY(cudaSetDevice(s.device));
        int rc;
        if ((rc = mg_stage(g, s, (int)k, &s.x0, x0, g->n))) return rc;
        if ((rc = gomp_batch_warm_start_x(s.h, s.x0))) return rc;
    }
    return GOMP_OK;
}

int gomp_mg_solve(gomp_mg *g) {
    if (!g) return fail(GOMP_ERR_INVALID_ARGUMENT, "null handle");
    for (MgShard &s : g->sh)
        if (s.count)
            if (int rc = gomp_batch_solve(s.h)) return rc;
    return GOMP_OK;
}

int gomp_mg_get(gomp_mg *g, double *x, double *y, int32_t *exit_code, int32_t *iters, int32_t *rho_updates) {
    if (!g) return fail(GOMP_ERR_INVALID_ARGUMENT, "null handle");
    const size_t G = g->sh.size();
    const bool want_ints = exit_code || iters || rho_updates;
    // queue every shard's copies first, then wait: the shards' transfers overlap
    for (size_t k = 0; k < G; ++k) {
        MgShard &s = g->sh[k];
        if (s.count == 0) continue;
        CUDA_TRY(cudaSetDevice(s.device));
        const double *xo = nullptr;
        gomp_batch_device_views(s.h, nullptr, nullptr, nullptr, &xo, nullptr);
        int rc;
        if (x && (rc = mg_rows_out(g, s, (int)k, x, xo, g->n))) return rc;
        if (y) {
            if (!s.yo) CUDA_TRY(cudaMalloc((void **)&s.yo, (size_t)s.count * (size_t)g->m * sizeof(double)));
            if ((rc = gomp_batch_get_async(s.h, nullptr, s.yo, nullptr, nullptr, nullptr))) return rc;
            if ((rc = mg_rows_out(g, s, (int)k, y, s.yo, g->m))) return rc;
        }
        if (want_ints) {
            if (!s.ints) CUDA_TRY(cudaMalloc((void **)&s.ints, (size_t)s.count * 3 * sizeof(int32_t)));
            if ((rc = gomp_batch_get_async(s.h, nullptr, nullptr, s.ints, s.ints + s.count, s.ints + 2 * s.count))) return rc;
            s.h_ints.resize((size_t)s.count * 3);
            CUDA_TRY(cudaMemcpyAsync(s.h_ints.data(), s.ints, (size_t)s.count * 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, s.stream));
        }
    }
    for (MgShard &s : g->sh)
        if (s.count)
            if (int rc = gomp_batch_sync(s.h)) return rc;
    if (want_ints) {
        // 4-byte rows: interleaved on the host (one strided copy per row would be slower than the solve)
        int32_t *dst[3] = {exit_code, iters, rho_updates};
        for (int a = 0; a < 3; ++a) {
            if (!dst[a]) continue;
            std::vector<int32_t> tmp;
            int32_t *out = dst[a];
            const bool dev = is_device_pointer(out);
            if (dev) {
                tmp.resize((size_t)g->batch);
                out = tmp.data();
            }
            for (size_t k = 0; k < G; ++k)
                for (int64_t i = 0; i < g->sh[k].count; ++i) out[(size_t)i * G + k] = g->sh[k].h_ints[(size_t)a * g->sh[k].count + i];
            if (dev) CUDA_TRY(cudaMemcpy(dst[a], tmp.data(), (size_t)g->batch * sizeof(int32_t), cudaMemcpyDefault));
        }
    }
    return GOMP_OK;
}

int gomp_mg_shard(gomp_mg *g, int32_t k, gomp_batch **handle, int32_t *device, int64_t *count) {
    if (!g || k < 0 || (size_t)k >= g->sh.size()) return fail(GOMP_ERR_INVALID_ARGUMENT, "bad shard index");
    if (handle) *handle = g->sh[(size_t)k].h;
    if (device) *device = g->sh[(size_t)k].device;
    if (count) *count = g->sh[(size_t)k].count;
    return GOMP_OK;
}

int gomp_mg_num_shards(gomp_mg *g) { return g ? (int)g->sh.size() : 0; }

} // extern "C"
