// gomp_stream.h -- interface of the long-horizon back end (gomp_stream.cu): per-CTA workspace layout in
// global memory, shared-memory carve-up and the launch entry points used by gomp_cuda.cu.
#ifndef GOMP_STREAM_H
#define GOMP_STREAM_H

#include "gomp_kernel.cuh"

namespace gomp_dev {

constexpr int kStreamRing = 8;      // generic path: factor slots (one stage each) in flight per substitution warp
constexpr int kStreamGroup = 8;     // fast path: stages per bulk copy
constexpr int kStreamGroupRing = 3; // fast path: groups in flight per substitution warp
constexpr int kStreamThreads = 512; // threads per CTA of the streamed kernel

struct DevStream {
    double *ws;       // [grid][stride] per-CTA workspace
    long long stride; // doubles, multiple of 2 (16-byte alignment of every region for the bulk copies)
    // region offsets inside one workspace (doubles): scaled A values, scaled triu(P), w = rho z - y / scratch,
    // rho per row, pivot-block inverses [T][bs][bs], ELL off-diagonal blocks, column scratch, scaled bounds
    long long wA, wP, wW, wRho, wDinv, wOff, wT, wLo, wHi;
    // shared memory (byte offsets)
    int oRHS, oRing, oPub, oDprev, oOffBuf, oV, oRed, oBar;
    const uint32_t *row_ell; // [roww][m_used] (pos << 16 | value index), 0xFFFFFFFF = none
    const uint16_t *col_ell; // [colw][n_pad] compact row of the j-th entry of a column
    int roww, colw;
    int slot_doubles; // bs*bs + bs*ow, one ring slot
    int ring_doubles; // ring size per substitution warp
};

cudaError_t prepare_stream(int bs, size_t smem_bytes, int *ctas_per_sm);
cudaError_t launch_stream(const DevSym &S, const DevBatch &B, const DevSettings &st, const DevStream &X, int grid, size_t smem_bytes,
                          cudaStream_t stream);

} // namespace gomp_dev
#endif
