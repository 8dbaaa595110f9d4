// gomp_duo.h -- interface of the two-slot back end (gomp_duo.cu, GOMP_LINSYS_BLOCKTRI_TMEM): shared-memory
// carve-up, owner-thread ELL views and the launch entry points used by gomp_cuda.cu.
#ifndef GOMP_DUO_H
#define GOMP_DUO_H

#include "gomp_kernel.cuh"

namespace gomp_dev {

constexpr int kDuoSlots = 2;                         // QPs resident per CTA (one CTA per SM)
constexpr int kDuoChainThreads = 64;                 // two substitution warps per slot (down / up elimination chain)
#ifndef GOMP_DUO_OWNERS
#define GOMP_DUO_OWNERS 192
#endif
constexpr int kDuoOwners = GOMP_DUO_OWNERS;          // row / variable owner threads per slot (6 warps: 16 warps per CTA, 128 registers per thread)
constexpr int kDuoSlotThreads = kDuoChainThreads + kDuoOwners;
constexpr int kDuoThreads = kDuoSlots * kDuoSlotThreads; // 512: warps 0..5 owners of slot 0, 6..11 owners of slot 1, 12..15 substitution (one per TMEM lane quarter)
constexpr int kDuoMaxRpt = 10, kDuoMaxCpt = 6;
constexpr int kDuoTmemCols = 512;                    // the whole tensor memory of the SM (one CTA per SM)

struct DevDuo {
    int rpt, cpt;
    int rlen[kDuoMaxRpt], rbase[kDuoMaxRpt]; // ELL-by-slot row view: entries of row slot j / byte offset of its first entry line
    int clen[kDuoMaxCpt], cbase[kDuoMaxCpt]; // ... column view of A
    int plen[kDuoMaxCpt], pbase[kDuoMaxCpt]; // ... full-symmetric view of P
    const uint32_t *row_ell, *col_ell, *pf_ell; // global copies ([line][owner]); row / col are staged into shared memory
    const uint16_t *col_pos;                    // [cpt][owner] position owned, 0xFFFF = none
    int n_row_ell, n_col_ell;                   // entries of the two staged views
    // assembly lists of S, ELL by owner (symbolic.h: build_assembly_ell); global memory
    const unsigned long long *as_con, *as_ocon;
    const uint32_t *as_meta, *as_ometa;
    int as_ept, as_maxc, as_opt, as_maxco;
    // shared memory (bytes): per-kernel pattern views, then one region block per slot
    int oRowEll, oColEll;
    int slot_base, slot_bytes;
    // ... offsets inside a slot block
    int oA, oP, oW, oRHS, oLo, oHi, oOff, oV, oScr, oRing, oRed, oCtl;
};

// bytes of factorisation scratch (previous inverse of both chains + pivot-column exchange) and of the S staging ring
inline size_t duo_scratch_bytes(int bs) { return 8 * ((size_t)2 * bs * bs + 64); }
inline size_t duo_ring_bytes(int bs) { return 8 * (size_t)2 * (bs * (bs + 1) / 2 + 1); }

cudaError_t prepare_duo(int rpt, int cpt, int bs, size_t smem_bytes, int *ctas_per_sm);
cudaError_t launch_duo(int rpt, int cpt, int bs, const DevSym &S, const DevBatch &B, const DevSettings &st, const DevDuo &X, int grid,
                       size_t smem_bytes, cudaStream_t stream);
bool duo_variant_exists(int rpt, int cpt, int bs);

} // namespace gomp_dev
#endif
