// gomp_stream.h -- interface of the long-horizon back end (gomp_stream.cu): per-CTA workspace layout in
// global memory, shared-memory carve-up and the launch entry points used by gomp_cuda.cu.
#ifndef GOMP_STREAM_H
#define GOMP_STREAM_H

#include "gomp_kernel.cuh"

namespace gomp_dev {

constexpr int kStreamRing = 8;      // generic path: factor slots (one stage each) in flight per substitution warp
constexpr int kStreamGroup = 8;     // fast path: stages per bulk copy (long horizons)
constexpr int kStreamGroupRing = 3; // fast path: groups in flight per substitution warp (long horizons)
constexpr int kStreamSmallThreads = 256; // small-QP mode: threads per CTA (two CTAs per SM)
constexpr int kStreamSmallGroup = 5, kStreamSmallGroupRing = 2; // ... and its ring geometry
constexpr int kStreamThreads = 512; // threads per CTA of the streamed kernel

struct DevStream {
    double *ws;       // [grid][stride] per-CTA workspace
    long long stride; // doubles, multiple of 2 (16-byte alignment of every region for the bulk copies)
    // region offsets inside one workspace (doubles): scaled A values, scaled triu(P), w = rho z - y / scratch,
    // rho per row, pivot-block inverses [T][bs][bs], ELL off-diagonal blocks, column scratch, scaled bounds
    long long wA, wP, wW, wRho, wDinv, wOff, wT, wLo, wHi;
    // shared memory (byte offsets)
    int oRHS, oRing, oPub, oDprev, oOffBuf, oV, oMid, oRed, oBar;
    const uint32_t *row_ell; // [roww][m_used] (pos << 16 | value index), 0xFFFFFFFF = none
    const uint16_t *col_ell; // [colw][n_pad] compact row of the j-th entry of a column
    int roww, colw;
    unsigned char rlen_blk[256]; // longest row of rows [256 k, 256 k + 256) of the compact (longest-first) row order; block 255 covers the rest
    int gen_ring;      // generic path: ring slots (one stage each) per substitution warp, <= kStreamRing
    int grp, grp_ring; // fast path: stages per bulk copy, groups in flight per substitution warp
    int e_scratch;     // ... and the row factors E behind them (offset into the ring in doubles, 0 = they stay in the workspace)
    int p_scratch;     // the P side of the Ruiz scaling fits the idle right-hand side / ring regions (shared memory)
    int small;         // small-QP mode: A values, w and the CSR / CSC pattern arrays in shared memory, 256 threads per CTA
    int oA, oW, oRowEnt, oRowRange, oColRange, oCscRow; // ... their byte offsets
    // Ruiz equilibration done beforehand by a cluster of CTAs per QP (gomp_ruiz_cluster_kernel; NULL = in-kernel): per-QP buffers
    // [batch][pre_stride] with the scaled A values, triu(P), bounds, the pass's column / row factors, the cluster's reduction
    // slots and the cost scale (offsets in doubles)
    double *pre;
    long long pre_stride, pA, pP, pLo, pHi, pT, pW, pRed, pC;
    int slot_doubles; // bs*bs + bs*ow, one ring slot
    int ring_doubles; // ring size per substitution warp
};

cudaError_t prepare_stream(int bs, bool small, size_t smem_bytes, int *ctas_per_sm);
// grid CTAs in clusters of `cluster_size` (1: plain launch; > 1: one QP per cluster, 512-thread kernels only)
cudaError_t launch_stream(const DevSym &S, const DevBatch &B, const DevSettings &st, const DevStream &X, int grid, int cluster_size,
                          size_t smem_bytes, cudaStream_t stream);
cudaError_t stream_max_clusters(int bs, int cluster_size, size_t smem_bytes, int *n);
// the cluster kernel: `clusters` clusters of `cluster_size` CTAs (1, 2, 4 or 8), QPs dealt round-robin to the clusters
cudaError_t launch_prescale(const DevSym &S, const DevBatch &B, const DevSettings &st, const DevStream &X, int clusters, int cluster_size,
                            cudaStream_t stream);

} // namespace gomp_dev
#endif
