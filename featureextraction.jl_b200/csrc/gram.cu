// gram.cu -- kernel-matrix construction: K = map(A B') with the kernel map, zero padding and
// per-tile row sums fused behind the contraction.  Replaces MLKernels.kernelmatrix as called at
// /root/reference/src/kpca.jl:51 (symmetric) and kpca.jl:65 (cross).
//
// One persistent CTA per SM, warp-specialised (512 threads, setmaxnreg 200 / 88 / 24):
//   producer (warp 12, one lane): per 8-feature step TWO tensor-map bulk copies (TMA, SASS UTMALDG),
//     boxes of {132 observations x 8 features} of the row and of the column operand -- the box is 4
//     elements wider than the 128 it feeds, which makes the shared-memory row pitch 132 doubles and
//     the LDS.64 fragment reads conflict free (t*132 + g distinct mod 16) -- into a 4-stage ring
//     guarded by mbarrier full/empty pairs; it runs ahead across tile boundaries.
//   8 MMA warps (2 x 4, 64 x 32 register tile each, DMMA.8x8x4 = mma.sync f64).  When a 128 x 128
//     tile is done they drop the RAW dot products into a 128 KB staging tile in shared memory (the
//     layout a 128-byte-swizzled TMA store expects) and start the next tile at once.  The kernel map
//     of tile T (branch-free exp / templated powers / tanh, in place in the staging tile) is
//     SOFTWARE-PIPELINED INTO THE MAIN LOOP OF TILE T+1: each k-step's 64 DMMAs share a basic block
//     with the map of a few staged elements, so the map's FP64 operations fill the issue gaps of
//     the DMMA stream instead of stopping it.  (DMMA and DFMA share one FP64 datapath.  Measured
//     this round: a map that runs as a burst after each tile -- round 1, inside the MMA warps --
//     leaves the datapath idle while both warps of a scheduler are in their bursts, 77-80 %; a map
//     in separate epilogue warps is starved by the DMMA stream, the MMA warps end up waiting for
//     the staging tile, 73-76 %.)
//   4 store warps: when the MMA warps have mapped a tile, ONE tensor-map bulk store (UTMASTG) sends
//     it to K; mirrored tiles are also written transposed, and the per-tile row / column sums that
//     feed the centring are formed here.
//
// Symmetry.  K is symmetric when A is a row block of B.  sym = 1 (one rank, or no peer access):
// inside this rank's square diagonal block only the tiles on / right of the diagonal are computed
// and the strictly-right ones are mirrored: the epilogue thread that owns tile row r writes its
// mapped values down column r of the transposed tile, so a warp's store instruction covers 32
// consecutive doubles (256 contiguous bytes) of one row of the mirrored tile.  sym = 2 (several
// ranks, peer arenas mapped through CUDA IPC): the off-diagonal blocks are shared out too -- of
// each pair of blocks (r,c) / (c,r) exactly one rank computes every tile and stores the transposed
// values, and the column sums that are the peer's row-sum partials, STRAIGHT INTO THE PEER'S K
// over NVLink (st.global to the IPC-mapped address), so every rank executes ~half of its row
// block and no separate exchange step exists.
#include <cuda.h>
#include <stdlib.h>
#include "fx_kernels.cuh"
#include "gram_tiles.cuh"

int fx_launch_gram_v3(fx_ctx* ctx, const GramArgs& a, int variant);      // gram_v3.cu

namespace {

constexpr int BM = FX_GRAM_BM, BN = FX_GRAM_BN, BK = 8, NST = 4;
constexpr int PITCH = 132;                                   // doubles per staged operand row
constexpr int SLAB_BYTES = BK * PITCH * 8;                   // 8448 (multiple of 128)
constexpr int STAGE_BYTES = 2 * SLAB_BYTES;
constexpr int ACC_BYTES = BM * BN * 8;                       // 131072: the staging tile
constexpr int OFF_STAGES = ACC_BYTES;
constexpr int OFF_COLW = OFF_STAGES + NST * STAGE_BYTES;     // [8 warps][32] pre-scaled column norms of the staged tile
constexpr int OFF_RSUM = OFF_COLW + 8 * 32 * 8;              // [8 warps][64] row sums of each warp's 64 x 32 block
constexpr int OFF_BARS = OFF_RSUM + 8 * 64 * 8;
constexpr int SMEM_BYTES = OFF_BARS + 128 + 1024;            // + slack for the 1024-byte alignment
constexpr int NTHREADS = 512;                                // 8 MMA + 4 store + 4 producer-group warps
constexpr int ST_THREADS = 128, ST_TID0 = 256;
constexpr int UNITS = 32;                                    // 16-byte units (2 elements) per lane per tile
constexpr uint64_t L2_EVICT_FIRST = 0x12F0000000000000ull;
constexpr uint64_t L2_EVICT_LAST = 0x14F0000000000000ull;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

struct GramParams {
    const double* A; long lda;
    const double* B; long ldb;
    const double* na; const double* nb;
    double* K; long ldk;
    double* rowpart; long rows_pad;
    long rows_valid, cols_valid, diag_shift;
    int kt;
    int tiles_x, tiles_y; long tiles_total;
    int sym;                          // 0 / 1 / 2, see the header comment
    int rank, world;                  // sym == 2: this rank's block index and the number of blocks
    double* peerK[16];                // sym == 2: K of rank c (own entry = K)
    double* peerRowpart[16];          //           rowpart of rank c (or null)
    KernelMap km;
};

// tile_role() / tile_coords(): gram_tiles.cuh (host-testable)

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar,
                                            uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(hint)
        : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, const void* src, int c0, int c1, int c2,
                                             uint64_t hint) {
    asm volatile(
        "cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%1, %2, %3}], [%4], %5;" ::"l"(
            reinterpret_cast<uint64_t>(tm)),
        "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(src)), "l"(hint)
        : "memory");
}
__device__ __forceinline__ void bar_sync_store() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// Byte offset of the 16-byte unit that holds columns (2u', 2u'+1) of chunk `c` in row `r` of the
// staging tile: [chunk of 16 columns][row][128 bytes], the eight 16-byte units of a row XOR-swizzled
// with (row & 7) -- the shared-memory image of a {16, 128, 8} box of a SWIZZLE_128B tensor map.
__device__ __forceinline__ int stage_off(int r, int c, int u) { return c * (BM * 128) + r * 128 + ((u ^ (r & 7)) << 4); }

// Map of one element; for the Gaussian kernel (KID 0) the norms arrive pre-scaled by -gamma.
template <int KID>
__device__ __forceinline__ double gram_map(const KernelMap& km, double dot, double ni, double nj, bool same) {
    if (KID == 0) return gaussian_map_prescaled(2.0 * km.p0, dot, ni, nj, same);
    return kernel_map_t<KID>(km, dot, ni, nj, same);
}

// Geometry of the in-place map: MMA warp (wm, wn) maps the 64 x 32 block it produced.  Lane L owns
// rows wm*64 + L and wm*64 + L + 32 and walks the 16 units (2 chunks x 8) of each: unit q = 0..31 is
// row L + 32 (q >> 4), chunk wn*2 + ((q >> 3) & 1), unit q & 7.  Lanes = consecutive rows, so the
// swizzle makes every quarter-warp hit 8 distinct 16-byte units: conflict free.
struct PrevTile {
    long i0, j0;            // coordinates of the tile that sits (raw) in the staging tile
    int q;                  // next unit to map (UNITS = done)
    bool valid, edge;
    double ni0, ni1;        // pre-scaled norms of this lane's two rows (Gaussian)
    double rs, rs0;         // running row sum of the mapped values (this warp's 32 columns); rs0 = first row's
                            // (scalars, no arrays: a run-time index would put the struct in local memory)
};

template <int KID>
__device__ __forceinline__ void map_unit(const GramParams& p, unsigned char* S, const double* colw, int wm, int wn,
                                         int lane, PrevTile& pv, int q) {
    const int rsel = q >> 4, c = wn * 2 + ((q >> 3) & 1), u = q & 7;
    double2* ptr = reinterpret_cast<double2*>(S + stage_off(wm * 64 + lane + 32 * rsel, c, u));
    const double2 dv = *ptr;
    double n0 = 0.0, n1 = 0.0;
    if (KID == 0) {
        const double2 nn = *reinterpret_cast<const double2*>(colw + (q & 15) * 2);         // broadcast read
        n0 = nn.x; n1 = nn.y;
    }
    const double ni = rsel ? pv.ni1 : pv.ni0;
    const double v0 = gram_map<KID>(p.km, dv.x, ni, n0, false);
    const double v1 = gram_map<KID>(p.km, dv.y, ni, n1, false);
    *ptr = make_double2(v0, v1);
    if (q == 16) { pv.rs0 = pv.rs; pv.rs = 0.0; }              // second row starts
    pv.rs += v0 + v1;
}

// EDGE tiles (ragged borders or touching the diagonal): validity masks and the same-point test;
// mapped in one go before the main loop of the next tile (they are O(n / 128) of O((n / 128)^2) tiles).
template <int KID>
__device__ __forceinline__ void map_edge_tile(const GramParams& p, unsigned char* S, const double* colw, int wm, int wn,
                                           int lane, PrevTile& pv) {
    for (int q = pv.q; q < UNITS; ++q) {
        const int rsel = q >> 4, c = wn * 2 + ((q >> 3) & 1), u = q & 7;
        const int row = wm * 64 + lane + 32 * rsel;
        double2* ptr = reinterpret_cast<double2*>(S + stage_off(row, c, u));
        const double2 dv = *ptr;
        double n0 = 0.0, n1 = 0.0;
        if (KID == 0) { n0 = colw[(q & 15) * 2]; n1 = colw[(q & 15) * 2 + 1]; }
        const long i = pv.i0 + row, j = pv.j0 + c * 16 + u * 2;
        const long ig = (p.diag_shift >= 0) ? i + p.diag_shift : -1;
        const bool iv = i < p.rows_valid;
        const double ni = rsel ? pv.ni1 : pv.ni0;
        double v0 = gram_map<KID>(p.km, dv.x, ni, n0, ig == j);
        double v1 = gram_map<KID>(p.km, dv.y, ni, n1, ig == j + 1);
        v0 = (iv && j < p.cols_valid) ? v0 : 0.0;
        v1 = (iv && j + 1 < p.cols_valid) ? v1 : 0.0;
        *ptr = make_double2(v0, v1);
        if (q == 16) { pv.rs0 = pv.rs; pv.rs = 0.0; }
        pv.rs += v0 + v1;
    }
    pv.q = UNITS;
}

// UPS = staged units mapped per k-step while a previous tile is pending (chosen by the launcher so
// that the map is finished after about half of the k-steps: 32 / UPS <= kt / 2 where possible).
template <int KID, int UPS>
__global__ void __launch_bounds__(NTHREADS, 1)
gram_ws_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmK, const GramParams p) {
    extern __shared__ unsigned char smem_dyn[];
    // 1024-byte alignment by pointer arithmetic on the __shared__ array itself (a round trip through
    // an integer makes the compiler lose the address space and emit generic LD / ST)
    unsigned char* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
    unsigned char* S = base;                                       // staging tile (1024-byte aligned)
    uint64_t* full = reinterpret_cast<uint64_t*>(base + OFF_BARS);
    uint64_t* empty = full + NST;
    uint64_t* mapped = empty + NST;                                // staging tile holds K values (8 MMA warps arrive)
    uint64_t* acc_empty = mapped + 1;                              // staging tile has been stored (store warps arrive)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KT = p.kt;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); }
        mbar_init(mapped, 8);
        mbar_init(acc_empty, 1);
        fence_mbar_init();
    }
    __syncthreads();

    if (warp >= 12) {
        // ------------------------------ producer warpgroup ------------------------------
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        if (warp != 12 || lane != 0) return;
        int it = 0;
        for (long tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x) {
            long i0, j0; int tx;
            tile_coords(p, tile, i0, j0, tx);
            int peer;
            if (tile_role(p, (int)(i0 / BM), tx, peer) == 0) continue;      // produced by a mirror store
            for (int kt = 0; kt < KT; ++kt, ++it) {
                const int s = it & (NST - 1);
                mbar_wait(&empty[s], ((it / NST) & 1) ^ 1);       // passes at once on first use
                unsigned char* st = base + OFF_STAGES + s * STAGE_BYTES;
                mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
                tma_load_2d(st, &tmA, (int)i0, kt * BK, &full[s], L2_EVICT_LAST);
                tma_load_2d(st + SLAB_BYTES, &tmB, (int)j0, kt * BK, &full[s], L2_EVICT_LAST);
            }
        }
        return;
    }

    if (warp >= 8) {
        // ------------------------------ store warpgroup ------------------------------
        asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
        const int e = tid - ST_TID0;                               // tile row (pass 1) / tile column (pass 2)
        const double* rsum_w = reinterpret_cast<const double*>(base + OFF_RSUM);
        int tcount = 0;
        for (long tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x) {
            long i0, j0; int tx;
            tile_coords(p, tile, i0, j0, tx);
            int peer;
            const int role = tile_role(p, (int)(i0 / BM), tx, peer);
            if (role == 0) continue;
            mbar_wait(mapped, tcount & 1);                         // the MMA warps finished the map of this tile
            if (e == 0) {
                tma_store_3d(&tmK, S, 0, (int)i0, (int)(j0 / 16), L2_EVICT_FIRST);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            if (p.rowpart) {
                // row-sum partials, one slot per 64 columns: the four column blocks of the MMA warps, pairwise
                const int wm = e >> 6, r = e & 63;
                const double* rw = rsum_w + (wm * 4) * 64 + r;
                p.rowpart[(long)(2 * tx) * p.rows_pad + i0 + e] = rw[0] + rw[64];
                p.rowpart[(long)(2 * tx + 1) * p.rows_pad + i0 + e] = rw[128] + rw[192];
            }
            if (role == 2) {
                double* Kmir = (p.sym == 2) ? p.peerK[peer] : p.K;
                double* rpmir = (p.sym == 2) ? p.peerRowpart[peer] : p.rowpart;
                const long mir_row0 = (p.sym == 2) ? (long)peer * p.rows_pad : p.diag_shift;
                // transposed tile: thread e walks row e of the staging tile and writes column
                // (i0 + e) of the owner's rows j0 - mir_row0 + ..: lanes = consecutive columns, so a
                // warp's store covers 256 contiguous bytes of one mirrored row
                double* mir = Kmir + (j0 - mir_row0) * p.ldk + (i0 + e + p.diag_shift);
#pragma unroll 1
                for (int c = 0; c < BN / 16; ++c) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const double2 v = *reinterpret_cast<const double2*>(S + stage_off(e, c, u));
                        const int col = c * 16 + u * 2;
                        __stcs(mir + (long)col * p.ldk, v.x);
                        __stcs(mir + (long)(col + 1) * p.ldk, v.y);
                    }
                }
                if (rpmir) {
                    // column sums = row-sum partials of the mirrored rows; thread e sums column e over
                    // the tile rows [0, 64) and [64, 128): columns [i0, +64), [i0 + 64, +64) of the owner
                    const int c = e >> 4, u = (e & 15) >> 1, hb = (e & 1) * 8;
                    double cs0 = 0.0, cs1 = 0.0;
#pragma unroll 8
                    for (int r = 0; r < 64; ++r) {
                        cs0 += *reinterpret_cast<const double*>(S + stage_off(r, c, u) + hb);
                        cs1 += *reinterpret_cast<const double*>(S + stage_off(r + 64, c, u) + hb);
                    }
                    const long slot = (i0 + p.diag_shift) / 64;
                    rpmir[slot * p.rows_pad + (j0 - mir_row0) + e] = cs0;
                    rpmir[(slot + 1) * p.rows_pad + (j0 - mir_row0) + e] = cs1;
                }
            }
            if (e == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            bar_sync_store();                                      // store engine and all readers are done with S
            if (e == 0) mbar_arrive(acc_empty);
            ++tcount;
        }
        if (e == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        return;
    }

    // -------------------------------- MMA warps --------------------------------
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    const int g = lane >> 2, t = lane & 3;
    const int wm = warp >> 2, wn = warp & 3;
    double* colw = reinterpret_cast<double*>(base + OFF_COLW) + warp * 32;
    double* rsum_w = reinterpret_cast<double*>(base + OFF_RSUM) + warp * 64;
    PrevTile pv;
    pv.valid = false; pv.edge = false; pv.q = UNITS; pv.i0 = pv.j0 = 0;
    pv.ni0 = pv.ni1 = pv.rs = pv.rs0 = 0.0;
    int it = 0, tcount = 0;

    // signals that the staged tile is mapped: row sums out, staging writes ordered before the TMA store
#define FX_GRAM_FINISH_PREV()                                                                                    \
    {                                                                                                            \
        rsum_w[lane] = pv.rs0;                                                                                   \
        rsum_w[lane + 32] = pv.rs;                                                                               \
        fence_proxy_async();                                                                                     \
        __syncwarp();                                                                                            \
        if (lane == 0) mbar_arrive(mapped);                                                                      \
        pv.valid = false;                                                                                        \
    }

    for (long tile = blockIdx.x; tile < p.tiles_total; tile += gridDim.x) {
        long i0, j0; int tx;
        tile_coords(p, tile, i0, j0, tx);
        int peer;
        if (tile_role(p, (int)(i0 / BM), tx, peer) == 0) continue;
        if (pv.valid && pv.edge) map_edge_tile<KID>(p, S, colw, wm, wn, lane, pv);   // before the accumulators are live
        double acc[8][4][2];
#pragma unroll
        for (int mt = 0; mt < 8; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

        // one k-step: 64 DMMAs on stage (it mod NST)
#define FX_GRAM_KSTEP()                                                                                          \
        {                                                                                                        \
            const int s = it & (NST - 1);                                                                        \
            mbar_wait(&full[s], (it / NST) & 1);                                                                 \
            const double* As = reinterpret_cast<const double*>(base + OFF_STAGES + s * STAGE_BYTES) + wm * 64 + g; \
            const double* Bs = reinterpret_cast<const double*>(base + OFF_STAGES + s * STAGE_BYTES + SLAB_BYTES) + wn * 32 + g; \
            _Pragma("unroll")                                                                                    \
            for (int kk = 0; kk < BK / 4; ++kk) {                                                                \
                double a[8], b[4];                                                                               \
                const int kr = kk * 4 + t;                                                                       \
                _Pragma("unroll")                                                                                \
                for (int mt = 0; mt < 8; ++mt) a[mt] = As[kr * PITCH + mt * 8];                                  \
                _Pragma("unroll")                                                                                \
                for (int nt = 0; nt < 4; ++nt) b[nt] = Bs[kr * PITCH + nt * 8];                                  \
                _Pragma("unroll")                                                                                \
                for (int mt = 0; mt < 8; ++mt)                                                                   \
                    _Pragma("unroll")                                                                            \
                    for (int nt = 0; nt < 4; ++nt) dmma(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);           \
            }                                                                                                    \
        }
#define FX_GRAM_KSTEP_END()                                                                                      \
        {                                                                                                        \
            __syncwarp();                                                                                        \
            if (lane == 0) mbar_arrive(&empty[(it - 1) & (NST - 1)]);                                            \
        }

        int kt = 0;
        if (pv.valid) {
            // phase 1: k-steps that also map UPS staged units of the previous tile (same basic block
            // as the DMMAs: the scheduler interleaves the map's instructions with them)
            for (; kt < KT && pv.q < UNITS; ++kt) {
                FX_GRAM_KSTEP();
#pragma unroll
                for (int j = 0; j < UPS; ++j) map_unit<KID>(p, S, colw, wm, wn, lane, pv, pv.q + j);
                pv.q += UPS;
                ++it;
                FX_GRAM_KSTEP_END();
            }
            for (; pv.q < UNITS; ++pv.q) map_unit<KID>(p, S, colw, wm, wn, lane, pv, pv.q);   // fewer k-steps than units
            FX_GRAM_FINISH_PREV();
        }
        // phase 2: plain k-steps
        for (; kt < KT; ++kt) {
            FX_GRAM_KSTEP();
            ++it;
            FX_GRAM_KSTEP_END();
        }

        // drop the raw dot products into the staging tile (free once the store warps have sent the
        // previous tile) and go on with the next tile
        mbar_wait(acc_empty, (tcount & 1) ^ 1);                  // passes at once for the first tile
#pragma unroll
        for (int mt = 0; mt < 8; ++mt) {
            const int r = wm * 64 + mt * 8 + g;                  // r & 7 == g
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int c = wn * 2 + (nt >> 1), u = (nt & 1) * 4 + t;
                *reinterpret_cast<double2*>(S + stage_off(r, c, u)) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
            }
        }
        // state of the tile now staged: norms (pre-scaled by -gamma), edge flag
        pv.valid = true; pv.i0 = i0; pv.j0 = j0; pv.q = 0; pv.rs = pv.rs0 = 0.0;
        {
            const long ig0 = (p.diag_shift >= 0) ? i0 + p.diag_shift : -(1L << 40);
            pv.edge = (i0 + BM > p.rows_valid) || (j0 + BN > p.cols_valid) || (ig0 < j0 + BN && ig0 + BM > j0);
        }
        if (KID == 0) {
            pv.ni0 = -p.km.p0 * p.na[i0 + wm * 64 + lane];
            pv.ni1 = -p.km.p0 * p.na[i0 + wm * 64 + lane + 32];
            colw[lane] = -p.km.p0 * p.nb[j0 + wn * 32 + lane];
        }
        __syncwarp();                                            // staging block and colw: this warp only
        ++tcount;
    }
    // the last tile of this CTA has no successor to hide behind
    if (pv.valid) {
        if (pv.edge) map_edge_tile<KID>(p, S, colw, wm, wn, lane, pv);
        for (; pv.q < UNITS; ++pv.q) map_unit<KID>(p, S, colw, wm, wn, lane, pv, pv.q);
        FX_GRAM_FINISH_PREV();
    }
#undef FX_GRAM_FINISH_PREV
#undef FX_GRAM_KSTEP
#undef FX_GRAM_KSTEP_END
}

// ---- naive cross-check variant: one thread per output element -------------------------------
__global__ void gram_naive_kernel(const GramParams p, int d_pad) {
    const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long i = blockIdx.y;
    if (j >= p.ldk) return;
    double v = 0.0;
    if (i < p.rows_valid && j < p.cols_valid) {
        double dot = 0.0;
        for (int f = 0; f < d_pad; ++f) dot += p.A[(long)f * p.lda + i] * p.B[(long)f * p.ldb + j];
        const double ni = (p.km.id == FX_KERNEL_GAUSSIAN) ? p.na[i] : 0.0;
        const double nj = (p.km.id == FX_KERNEL_GAUSSIAN) ? p.nb[j] : 0.0;
        const long ig = (p.diag_shift >= 0) ? i + p.diag_shift : -1;
        v = kernel_map(p.km, dot, ni, nj, ig == j);
    }
    p.K[i * p.ldk + j] = v;
}
// naive per-tile row sums so both variants feed the same reduction
__global__ void gram_naive_rowpart_kernel(const double* K, long ldk, double* rowpart, long rows_pad) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long jt = blockIdx.y;
    if (i >= rows_pad) return;
    double s = 0.0;
    for (int c = 0; c < 64; ++c) s += K[i * ldk + jt * 64 + c];
    rowpart[jt * rows_pad + i] = s;
}

__global__ void rownorms_kernel(const double* P, long ld, long nobs, int d_pad, double* norms) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nobs) return;
    double s = 0.0;
    for (int f = 0; f < d_pad; ++f) {
        const double x = P[(long)f * ld + i];
        s = fma(x, x, s);
    }
    norms[i] = s;
}

bool use_v3() {
    static const bool v = []() { const char* e = getenv("FXB200_GRAM_V3"); return e && e[0] == '1'; }();
    return v;
}

template <int KID, int UPS>
int launch_ws2(fx_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmK,
               const GramParams& p, unsigned grid) {
    FX_SET_SMEM_ONCE(ctx, (gram_ws_kernel<KID, UPS>), SMEM_BYTES);
    gram_ws_kernel<KID, UPS><<<grid, NTHREADS, SMEM_BYTES, ctx->stream>>>(tmA, tmB, tmK, p);
    return 0;
}
template <int KID>
int launch_ws(fx_ctx* ctx, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmK,
              const GramParams& p, unsigned grid) {
    // units mapped per k-step: the map of the previous tile should be done after about half of the
    // k-steps, so that the store warps have the other half to send it out
    const int kt = p.kt;
    if (kt >= 64) return launch_ws2<KID, 1>(ctx, tmA, tmB, tmK, p, grid);
    if (kt >= 32) return launch_ws2<KID, 2>(ctx, tmA, tmB, tmK, p, grid);
    if (kt >= 16) return launch_ws2<KID, 4>(ctx, tmA, tmB, tmK, p, grid);
    return launch_ws2<KID, 8>(ctx, tmA, tmB, tmK, p, grid);
}

}  // namespace

// Number of per-tile row-sum partials (slots of `rowpart`) the Gram kernel writes per row.
int64_t fx_gram_rowpart_slots(int64_t ldk) { return ldk / 64; }

int fx_launch_rownorms(fx_ctx* ctx, const double* P, int64_t ld, int64_t nobs, int64_t d_pad,
                       double* norms) {
    if (nobs <= 0) return 0;
    rownorms_kernel<<<(unsigned)ceil_div(nobs, 256), 256, 0, ctx->stream>>>(P, ld, nobs, (int)d_pad, norms);
    ctx->launches++;
    FX_LAUNCH_CHECK();
    return 0;
}

int fx_launch_gram(fx_ctx* ctx, const GramArgs& a, int variant) {
    if (use_v3()) return fx_launch_gram_v3(ctx, a, variant);
    if (a.rows_pad % BM || a.ldk % BN || a.d_pad % 16 || a.lda % 2 || a.ldb % 2) {
        fx_set_error("gram: bad padding rows_pad=%lld ldk=%lld d_pad=%lld", (long long)a.rows_pad,
                     (long long)a.ldk, (long long)a.d_pad);
        return FX_ERR_CUDA;
    }
    GramParams p;
    p.A = a.A; p.lda = a.lda; p.B = a.B; p.ldb = a.ldb; p.na = a.na; p.nb = a.nb;
    p.K = a.K; p.ldk = a.ldk; p.rowpart = a.rowpart; p.rows_pad = a.rows_pad;
    p.rows_valid = a.rows_valid; p.cols_valid = a.cols_valid; p.diag_shift = a.diag_shift;
    p.kt = (int)(a.d_pad / BK);
    p.tiles_x = (int)(a.ldk / BN);
    p.tiles_y = (int)(a.rows_pad / BM);
    p.tiles_total = (long)(a.ldk / BN) * (a.rows_pad / BM);
    // whole symmetric matrix on one rank: only tiles on / right of the diagonal block are computed
    p.sym = (a.sym_source && a.diag_shift >= 0 && a.diag_shift % BM == 0 &&
             a.diag_shift + a.rows_pad <= a.ldk && !a.no_symmetry) ? 1 : 0;
    p.rank = 0; p.world = 1;
    for (int c = 0; c < 16; ++c) { p.peerK[c] = nullptr; p.peerRowpart[c] = nullptr; }
    if (p.sym && a.peer_world > 1) {
        // cross-rank symmetry: needs every rank's block to be rows_pad wide and all peers mapped
        if (a.peer_world > 16 || a.ldk != (int64_t)a.peer_world * a.rows_pad || a.diag_shift % a.rows_pad ||
            a.peer_rank != (int)(a.diag_shift / a.rows_pad) || variant != 0) {
            fx_set_error("gram: inconsistent peer layout (world %d, rank %d)", a.peer_world, a.peer_rank);
            return FX_ERR_CUDA;
        }
        p.sym = 2; p.world = a.peer_world; p.rank = a.peer_rank;
        for (int c = 0; c < a.peer_world; ++c) { p.peerK[c] = a.peerK[c]; p.peerRowpart[c] = a.peerRowpart[c]; }
    }
    p.km.id = a.kernel_id; p.km.p0 = a.kp[0]; p.km.p1 = a.kp[1]; p.km.p2 = a.kp[2];
    double tmp[3];
    if (fx_make_kernel_map(a.kernel_id, a.kp, &p.km.ipow, tmp)) {
        fx_set_error("gram: invalid kernel id / parameters");
        return -1;
    }
    if (a.rows_pad == 0 || a.ldk == 0) return 0;
    if (variant == 0) {
        PFN_encodeTiled enc = get_encode();
        if (!enc) {
            fx_set_error("gram: cuTensorMapEncodeTiled is not available from this driver");
            return FX_ERR_CUDA;
        }
        CUtensorMap tmA, tmB, tmK;
        auto operand_map = [&](CUtensorMap* tm, const double* ptr, int64_t nobs, int64_t ld) -> int {
            // feature-major operand: dim0 = observations (contiguous), dim1 = features
            cuuint64_t dims[2] = {(cuuint64_t)nobs, (cuuint64_t)a.d_pad};
            cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
            cuuint32_t box[2] = {PITCH, BK};
            cuuint32_t es[2] = {1, 1};
            CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)ptr, dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { fx_set_error("gram: operand tensor map rejected (CUresult %d)", (int)r); return FX_ERR_CUDA; }
            return 0;
        };
        FX_TRY(operand_map(&tmA, a.A, a.rows_pad, a.lda));
        FX_TRY(operand_map(&tmB, a.B, a.ldk, a.ldb));
        {
            cuuint64_t dims[3] = {16, (cuuint64_t)a.rows_pad, (cuuint64_t)(a.ldk / 16)};
            cuuint64_t strides[2] = {(cuuint64_t)a.ldk * 8, 128};
            cuuint32_t box[3] = {16, BM, BN / 16};
            cuuint32_t es[3] = {1, 1, 1};
            CUresult r = enc(&tmK, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)a.K, dims, strides, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) { fx_set_error("gram: tensor map for K rejected (CUresult %d)", (int)r); return FX_ERR_CUDA; }
        }
        const long want = ctx->sm_count > 0 ? ctx->sm_count : 148;
        const unsigned grid = (unsigned)(p.tiles_total < want ? p.tiles_total : want);
        switch (kernel_variant(p.km.id, p.km.ipow)) {
            case 0: FX_TRY(launch_ws<0>(ctx, tmA, tmB, tmK, p, grid)); break;
            case 1: FX_TRY(launch_ws<1>(ctx, tmA, tmB, tmK, p, grid)); break;
            case 2: FX_TRY(launch_ws<2>(ctx, tmA, tmB, tmK, p, grid)); break;
            case 3: FX_TRY(launch_ws<3>(ctx, tmA, tmB, tmK, p, grid)); break;
            case 4: FX_TRY(launch_ws<4>(ctx, tmA, tmB, tmK, p, grid)); break;
            case 5: FX_TRY(launch_ws<5>(ctx, tmA, tmB, tmK, p, grid)); break;
            default: FX_TRY(launch_ws<6>(ctx, tmA, tmB, tmK, p, grid)); break;
        }
        ctx->launches++;
    } else {
        dim3 grid((unsigned)ceil_div(a.ldk, 256), (unsigned)a.rows_pad);
        gram_naive_kernel<<<grid, 256, 0, ctx->stream>>>(p, (int)a.d_pad);
        ctx->launches++;
        if (a.rowpart) {
            dim3 g2((unsigned)ceil_div(a.rows_pad, 256), (unsigned)(a.ldk / 64));
            gram_naive_rowpart_kernel<<<g2, 256, 0, ctx->stream>>>(a.K, a.ldk, a.rowpart, a.rows_pad);
            ctx->launches++;
        }
    }
    FX_LAUNCH_CHECK();
    return 0;
}
