// Warp-chain passes of the fused 2D TM kernel: a pass of T = G*K time steps as a PIPELINE OF WARPS, fed by TMA.
//
// The register pipeline of fd2d_march.cu lets ONE warp carry a row through all T stages, so it keeps T+1 row sets in
// registers (255 registers at T = 6: eight warps per SM, two per scheduler) and every stage waits on the one before it:
// the kernel sits on its own latency roof (issue slots 42 % busy) exactly where it reaches the HBM roof.  The deep
// kernels of fd2d_deep.cu buy depth 8 by moving the accumulators into shared memory, which costs them ~20 B of
// shared-memory traffic per cell-update and leaves them at the same eight warps.
//
// Here a (strip, chunk) item belongs to a GROUP of G warps.  Warp g owns the K stages [g*K, (g+1)*K): it keeps K+1
// row sets in registers (84 registers of state at K = 2, so sixteen warps fit an SM), receives a row, runs its K
// stages -- the same packed stage march_stage_pk as the register pipeline, operation for operation -- and hands the row
// that leaves its last stage to warp g+1 through a small ring of shared-memory slots.  The G warps work on G different
// time levels of neighbouring rows at the same moment: the dependent chain of a row is spread over G schedulers' worth
// of issue slots instead of one.  Hand-off traffic is 6 arrays x 16 B per lane per K stages (12 B per cell-update at
// K = 2 counting both directions).
//
//   TMA  -- cp.async.bulk.tensor.2d, one elected lane --> staging ring of R-row x 128-column boxes, one box per array
//   warp 0   stages 0 .. K-1      --> queue 0 (QD slots of 6 x 512 B) -->
//   warp 1   stages K .. 2K-1     --> queue 1 --> ...
//   warp G-1 stages (G-1)K .. T-1 --> global stores (16 B per lane, the strip's inner columns and the chunk's rows only)
//
// Every hand-off is an mbarrier pair per slot (full: producer -> consumer, empty: consumer -> producer); the staging
// ring's full barriers are armed with the box bytes (expect_tx) and completed by the TMA unit.  Nothing in the kernel
// is a CTA-wide barrier after set-up.
//
// Row bookkeeping.  Stage 0 is fed the rows [r_begin, r_end) = [i0 - T, i1 + T) of the chunk.  A warp that has received
// its x-th row releases, after its K stages, the row it received K rows earlier; its first K releases would be the
// all-zero sets above the chunk and are dropped, so EVERY warp's x-th input is global row r_begin + x, warp g receives
// r_end - r_begin - g*K rows, and the last warp stores row r_begin + x - K.  Rows outside a warp's validity cone carry
// garbage exactly as in the single-warp pipeline; the rows stored, [i0, i1), depend on valid inputs only.
//
// Interior (FAST) items only: float, 4 columns per lane, lossless, no fused DFT.  Edges, PML, TFSF, source and halo
// exchange stay with k_careful2 (fd2d_deep.cu), which takes any depth.
#include "fd2d_march.cuh"

#include <cuda.h>          // CUtensorMap and its enums (types only: the encoder is resolved through the runtime)
#include <mutex>

namespace {

using namespace fdtd_march;

constexpr int CV = 4;                 // columns per lane
constexpr int CLB = CV * 4;           // bytes per lane per array row
constexpr int CROWB = 32 * CLB;       // bytes per array row of a warp (128 columns)
constexpr int NARR = 6;               // arrays that travel with a row: dz hx hy ihx ihy naz (ez is recomputed on arrival)

struct ChainMaps { CUtensorMap m[NARR]; };
enum { K_INTERIOR = 0, K_COL = 1, K_ROW = 2 };   // what a launch's items are: interior, (special strip, ordinary chunk), (ordinary strip, special chunk)   // in_dz in_hx in_hy in_ihx in_ihy naz: (rows, ny) float, box R x 128

template <int G_, int K_, int GROUPS_, int R_ = K_ + 1, int NSTAGE_ = 2, bool STS_PTX_ = false>
struct ChainShape {
    static constexpr bool STS_PTX = STS_PTX_;
    static constexpr int G = G_, K = K_, GROUPS = GROUPS_;
    static constexpr int T = G * K, NS = K + 1;
    static_assert(T == 8 || T == 12, "depths the host plans");
    // Ring geometry tied to the register rotation: a TMA box is NS rows x 128 columns of one array (one trip of the
    // unrolled row loop), the staging ring has two boxes per array, a queue has NS one-row slots -- so every slot index
    // is a compile-time constant of the unrolled loop and a barrier's parity is a bit of the trip counter.  (One-row
    // boxes in an NS-slot ring halve the staging memory but measured 23 % slower: 685 against 890 Gcell/s at
    // G = 2, K = 4 -- six 512-byte TMA operations and an elected issue per row instead of per trip.)
    // Smaller boxes (R_ < NS rows, NSTAGE_ of them) walk the staging ring with run-time counters instead: a few uniform
    // instructions per row buy back staging memory, i.e. room for more groups per SM.
    static constexpr int R = R_, NSTAGE = NSTAGE_, QD = NS;
    static constexpr bool STATIC_RING = (R == NS && NSTAGE == 2);
    static constexpr int HALO = ((T + CV - 1) / CV) * CV, W = 32 * CV, USE = W - 2 * HALO;
    static constexpr int BOX_B = R * CROWB;                   // one array of one box
    static constexpr int STAGE_B = NARR * BOX_B;              // one staging slot
    static constexpr int QSLOT_B = NARR * CROWB;              // one queue slot (one row)
    static constexpr int NBAR = NSTAGE + 2 * (G - 1) * QD;
    static constexpr int OFF_Q = NSTAGE * STAGE_B;
    static constexpr int OFF_BAR = OFF_Q + (G - 1) * QD * QSLOT_B;
    static constexpr int GROUP_SMEM = ((OFF_BAR + NBAR * 8 + 127) / 128) * 128;
    static constexpr int THREADS = GROUPS * G * 32;
    static constexpr int SMEM = GROUPS * GROUP_SMEM;
    static_assert(SMEM <= 227 * 1024, "the groups of a CTA must fit the shared memory of an SM");
};

// ---- mbarrier and TMA in PTX (sm_90+).  `leader`: the instruction is predicated on it (one lane), not branched around.
__device__ __forceinline__ unsigned smem_u32(const void *p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(const unsigned bar, const unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(const unsigned bar, const int leader) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.s32 p, %1, 0;\n"
        "@p mbarrier.arrive.shared::cta.b64 _, [%0];\n"
        "}\n" ::"r"(bar), "r"(leader) : "memory");
}
__device__ __forceinline__ void mbar_wait(const unsigned bar, const unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "CHAIN_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra CHAIN_WAIT;\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// one staging slot: arm its barrier with the bytes of the six boxes, then one box (R rows x 128 columns) per array
// global -> shared, completion counted in bytes on the barrier.  Called by the whole (converged) warp under a
// warp-uniform condition; elect.sync picks the one lane that issues, which lets the assembler keep every operand in
// uniform registers (a predicate derived from the lane id costs a waterfall loop per instruction).
__device__ __forceinline__ void tma_issue_boxes(const unsigned dst, const int box_bytes, const ChainMaps &maps, const unsigned bar,
                                                const int col, const int row) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "@p mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %9;\n"
        "@p cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%2, {%8, %10}], [%1];\n"
        "@p cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%11], [%3, {%8, %10}], [%1];\n"
        "@p cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%12], [%4, {%8, %10}], [%1];\n"
        "@p cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%13], [%5, {%8, %10}], [%1];\n"
        "@p cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%14], [%6, {%8, %10}], [%1];\n"
        "@p cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%15], [%7, {%8, %10}], [%1];\n"
        "}\n"
        ::"r"(dst), "r"(bar), "l"(&maps.m[0]), "l"(&maps.m[1]), "l"(&maps.m[2]), "l"(&maps.m[3]), "l"(&maps.m[4]), "l"(&maps.m[5]),
          "r"(col), "r"(NARR * box_bytes), "r"(row), "r"(dst + box_bytes), "r"(dst + 2 * box_bytes),
          "r"(dst + 3 * box_bytes), "r"(dst + 4 * box_bytes), "r"(dst + 5 * box_bytes)
        : "memory");
}

// Queue stores.  Values computed in register PAIRS are not aligned quads: the compiler fuses two adjacent 64-bit
// stores into one 128-bit store plus four moves.  PTX = true spells the 64-bit stores out (no moves, twice the store
// instructions, and volatile statements the scheduler cannot move the stage arithmetic across; no memory clobber: the
// stores only have to stay ordered with the barrier operations, which volatile gives).  Measured: the compiler's form wins
// at eight warps per SM (889 against 845 Gcell/s), the PTX form at twelve (885 against 848).
template <bool PTX>
__device__ __forceinline__ void sts22(void *dst, const float (&d)[CV]) {
    if constexpr (PTX) {
        const unsigned a = smem_u32(dst);
        asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(d[0]), "f"(d[1]));
        asm volatile("st.shared.v2.f32 [%0+8], {%1, %2};" ::"r"(a), "f"(d[2]), "f"(d[3]));
    } else {
        *reinterpret_cast<float2 *>(dst) = make_float2(d[0], d[1]);
        *reinterpret_cast<float2 *>(reinterpret_cast<unsigned char *>(dst) + 8) = make_float2(d[2], d[3]);
    }
}

// Column coefficients of a lane for the column variant: the five y-vectors of the PML at its four columns (identity
// outside the grid), gy2 already halved (gy2 * 0.5 is what the careful stage multiplies the curl by: (gx2*gy2)*0.5 with
// gx2 = 1), and the two cells whose update is masked: D of column 0 and H of column ny-1.
struct ColLane {
    float gy2h[CV], gy3[CV], fy1[CV], fy2[CV], fy3[CV];
    bool fix_d, fix_h;             // this lane owns column 0 (as its v = 0) / column ny-1 (as its v = CV-1)
};

// One stage on a strip whose COLUMNS carry PML coefficients while its rows are ordinary (gx2 = gx3 = fx2 = fx3 = 1,
// fx1 = 0): operation for operation march_stage<float, 4, 0, false> with the row coefficients at those values --
//   dz = (gy3*dz) + ((gy2*0.5)*curl);  ez = naz*dz
//   ihx += cm; ihy += cn;  hx = (fy3*hx) + (fy2*((0.5*cm) + (0*ihx)));  hy = hy - ((0.5*cn) + (fy1*ihy))
// -- in packed arithmetic.  Multiplications by a row coefficient that is exactly 1 are dropped (exact), 0*ihx is kept (it
// decides the sign of a zero sum).  The two masked cells are computed like the others and put back afterwards.
__device__ __forceinline__ void march_stage_pk_col(RowSet<float, CV> &A, RowSet<float, CV> &Hd, const ColLane &c, const float2 negzero) {
    constexpr unsigned FULL = 0xffffffffu;
    const float2 half2 = make_float2(0.5f, 0.5f), zero2 = make_float2(0.f, 0.f);
    // ---- D of the arriving row
    const float keep_dz = A.dz[0];
    const float hx_left = __shfl_up_sync(FULL, A.hx[CV - 1], 1);
#pragma unroll
    for (int v = 0; v < CV; v += 2) {
        const float2 a1 = pk_sub(make_float2(A.hy[v], A.hy[v + 1]), make_float2(Hd.hy[v], Hd.hy[v + 1]));
        const float2 a2 = pk_sub(a1, make_float2(A.hx[v], A.hx[v + 1]));
        const float2 curl = make_float2(a2.x + (v == 0 ? hx_left : A.hx[v == 0 ? 0 : v - 1]), a2.y + A.hx[v]);   // shifted pair: scalar
        const float2 dn = pk_add(pk_mul(make_float2(c.gy3[v], c.gy3[v + 1]), make_float2(A.dz[v], A.dz[v + 1]), negzero),
                                 pk_mul(make_float2(c.gy2h[v], c.gy2h[v + 1]), curl, negzero));
        A.dz[v] = dn.x; A.dz[v + 1] = dn.y;
    }
    if (c.fix_d) A.dz[0] = keep_dz;                   // column 0 has no D update
    // ---- E of the arriving row; the held row's Ez comes from its own trip
    float ezA[CV], ezH[CV];
#pragma unroll
    for (int v = 0; v < CV; v += 2) {
        const float2 a = pk_mul(make_float2(A.naz[v], A.naz[v + 1]), make_float2(A.dz[v], A.dz[v + 1]), negzero);
        A.ez[v] = a.x; A.ez[v + 1] = a.y;
        ezA[v] = a.x; ezA[v + 1] = a.y; ezH[v] = Hd.ez[v]; ezH[v + 1] = Hd.ez[v + 1];
    }
    // ---- H of the held row
    const float k_ihx = Hd.ihx[CV - 1], k_ihy = Hd.ihy[CV - 1], k_hx = Hd.hx[CV - 1], k_hy = Hd.hy[CV - 1];
    const float ez_right = __shfl_down_sync(FULL, ezH[0], 1);
#pragma unroll
    for (int v = 0; v < CV; v += 2) {
        const float2 e = make_float2(ezH[v], ezH[v + 1]);
        const float2 cm = make_float2(ezH[v] - ezH[v + 1], ezH[v + 1] - (v + 2 < CV ? ezH[v + 2 < CV ? v + 2 : v] : ez_right));   // shifted pair: scalar
        const float2 cn = pk_sub(e, make_float2(ezA[v], ezA[v + 1]));
        const float2 sx = pk_add(make_float2(Hd.ihx[v], Hd.ihx[v + 1]), cm);
        const float2 sy = pk_add(make_float2(Hd.ihy[v], Hd.ihy[v + 1]), cn);
        const float2 tx = pk_add(pk_mul(half2, cm, negzero), pk_mul(zero2, sx, negzero));
        const float2 ty = pk_add(pk_mul(half2, cn, negzero), pk_mul(make_float2(c.fy1[v], c.fy1[v + 1]), sy, negzero));
        const float2 hx2 = pk_add(pk_mul(make_float2(c.fy3[v], c.fy3[v + 1]), make_float2(Hd.hx[v], Hd.hx[v + 1]), negzero),
                                  pk_mul(make_float2(c.fy2[v], c.fy2[v + 1]), tx, negzero));
        const float2 hy2 = pk_sub(make_float2(Hd.hy[v], Hd.hy[v + 1]), ty);
        Hd.ihx[v] = sx.x; Hd.ihx[v + 1] = sx.y; Hd.ihy[v] = sy.x; Hd.ihy[v + 1] = sy.y;
        Hd.hx[v] = hx2.x; Hd.hx[v + 1] = hx2.y; Hd.hy[v] = hy2.x; Hd.hy[v + 1] = hy2.y;
    }
    if (c.fix_h) {                                    // column ny-1 has no H update
        Hd.ihx[CV - 1] = k_ihx; Hd.ihy[CV - 1] = k_ihy; Hd.hx[CV - 1] = k_hx; Hd.hy[CV - 1] = k_hy;
    }
}

// One stage on a chunk whose ROWS carry PML coefficients (or are the grid's first / last rows) while its columns are
// ordinary (gy2 = gy3 = fy2 = fy3 = 1, fy1 = 0): operation for operation march_stage<float, 4, 0, false> with the column
// coefficients at those values --
//   dz = (gx3*dz) + ((gx2*0.5)*curl)   where 1 <= rs < nx;   ez = naz*dz
//   ihx += cm; ihy += cn;  hx = hx + ((0.5*cm) + (fx1*ihx));  hy = (fx3*hy) - (fx2*((0.5*cn) + (0*ihy)))   where 0 <= hr <= nx-2
// -- in packed arithmetic; the row coefficients and the two row masks are warp-uniform.  rs: global row of the arriving
// row (the held row is rs - 1); rows outside the stored rows arrive as zeros (the tensor map fills them), exactly what
// the careful kernel's zero-filling fetch gives.
__device__ __forceinline__ void march_stage_pk_row(const MarchParams<float> &p, RowSet<float, CV> &A, RowSet<float, CV> &Hd,
                                                   const int rs, const float2 negzero) {
    constexpr unsigned FULL = 0xffffffffu;
    const float2 half2 = make_float2(0.5f, 0.5f), zero2 = make_float2(0.f, 0.f);
    const int hr = rs - 1;
    const int rd = min(max(rs, 0), p.nx - 1), rh = min(max(hr, 0), p.nx - 1);
    const float gx2h = __ldg(p.gx2 + rd) * 0.5f, gx3 = __ldg(p.gx3 + rd);       // (gx2 * gy2) * 0.5 with gy2 = 1
    const float fx1 = __ldg(p.fx1 + rh), fx2 = __ldg(p.fx2 + rh), fx3 = __ldg(p.fx3 + rh);
    const bool drow = rs >= 1 && rs < p.nx, hrow = hr >= 0 && hr <= p.nx - 2;
    // ---- D of the arriving row
    const float hx_left = __shfl_up_sync(FULL, A.hx[CV - 1], 1);
    if (drow) {
        const float2 g3 = make_float2(gx3, gx3), g2 = make_float2(gx2h, gx2h);
#pragma unroll
        for (int v = 0; v < CV; v += 2) {
            const float2 a1 = pk_sub(make_float2(A.hy[v], A.hy[v + 1]), make_float2(Hd.hy[v], Hd.hy[v + 1]));
            const float2 a2 = pk_sub(a1, make_float2(A.hx[v], A.hx[v + 1]));
            const float2 curl = make_float2(a2.x + (v == 0 ? hx_left : A.hx[v == 0 ? 0 : v - 1]), a2.y + A.hx[v]);   // shifted pair: scalar
            const float2 dn = pk_add(pk_mul(g3, make_float2(A.dz[v], A.dz[v + 1]), negzero), pk_mul(g2, curl, negzero));
            A.dz[v] = dn.x; A.dz[v + 1] = dn.y;
        }
    }
    // ---- E of the arriving row; the held row's Ez comes from its own trip
    float ezA[CV], ezH[CV];
#pragma unroll
    for (int v = 0; v < CV; v += 2) {
        const float2 a = pk_mul(make_float2(A.naz[v], A.naz[v + 1]), make_float2(A.dz[v], A.dz[v + 1]), negzero);
        A.ez[v] = a.x; A.ez[v + 1] = a.y;
        ezA[v] = a.x; ezA[v + 1] = a.y; ezH[v] = Hd.ez[v]; ezH[v + 1] = Hd.ez[v + 1];
    }
    // ---- H of the held row
    const float ez_right = __shfl_down_sync(FULL, ezH[0], 1);
    if (hrow) {
        const float2 f1 = make_float2(fx1, fx1), f2 = make_float2(fx2, fx2), f3 = make_float2(fx3, fx3);
#pragma unroll
        for (int v = 0; v < CV; v += 2) {
            const float2 e = make_float2(ezH[v], ezH[v + 1]);
            const float2 cm = make_float2(ezH[v] - ezH[v + 1], ezH[v + 1] - (v + 2 < CV ? ezH[v + 2 < CV ? v + 2 : v] : ez_right));   // shifted pair: scalar
            const float2 cn = pk_sub(e, make_float2(ezA[v], ezA[v + 1]));
            const float2 sx = pk_add(make_float2(Hd.ihx[v], Hd.ihx[v + 1]), cm);
            const float2 sy = pk_add(make_float2(Hd.ihy[v], Hd.ihy[v + 1]), cn);
            const float2 tx = pk_add(pk_mul(half2, cm, negzero), pk_mul(f1, sx, negzero));
            const float2 ty = pk_add(pk_mul(half2, cn, negzero), pk_mul(zero2, sy, negzero));
            const float2 hx2 = pk_add(make_float2(Hd.hx[v], Hd.hx[v + 1]), tx);
            const float2 hy2 = pk_sub(pk_mul(f3, make_float2(Hd.hy[v], Hd.hy[v + 1]), negzero), pk_mul(f2, ty, negzero));
            Hd.ihx[v] = sx.x; Hd.ihx[v + 1] = sx.y; Hd.ihy[v] = sy.x; Hd.ihy[v + 1] = sy.y;
            Hd.hx[v] = hx2.x; Hd.hx[v + 1] = hx2.y; Hd.hy[v] = hy2.x; Hd.hy[v + 1] = hy2.y;
        }
    }
}

// One warp of the chain.  FIRST: input from the TMA staging ring; else from the queue behind it.  LAST: output to global
// memory; else into the queue ahead.  Every warp runs the same number of trips: a warp hands on EVERY row that leaves
// its last stage, the all-zero sets of the first K sub-iterations included (zero rows stay zero through a stage), so
// the x-th input of warp g is global row r_begin + x - g*K and nothing in the loop depends on the warp's position.
template <typename Shape, bool FIRST, bool LAST, int KIND>
__device__ __forceinline__ void chain_body(const MarchParams<float> &p, const ChainMaps &maps, const int strip, const int i0,
                                           const int i1, const int lane, const int wg, unsigned char *const gsm) {
    constexpr int K = Shape::K, NSTAGE = Shape::NSTAGE;
    constexpr int T = Shape::T, NS = Shape::NS, HALO = Shape::HALO, W = Shape::W, USE = Shape::USE;
    static_assert(Shape::QD == NS, "queue slot indices follow the register rotation");
    constexpr bool SRING = Shape::STATIC_RING;
    constexpr int R = Shape::R;

    const int c0 = strip * USE - HALO;               // first column of the strip (halo included)
    const int jb = c0 + lane * CV;                   // first column of this lane
    constexpr bool COL = KIND == K_COL, ROW = KIND == K_ROW;
    const bool col_in = !COL || (jb >= 0 && jb + CV <= p.ny);     // (ordinary strips lie inside the grid)
    const bool col_store = col_in && (lane * CV >= HALO) && (lane * CV + CV <= W - HALO);
    ColLane cl;
    if (COL) {
#pragma unroll
        for (int v = 0; v < CV; ++v) { cl.gy2h[v] = 0.5f; cl.gy3[v] = cl.fy2[v] = cl.fy3[v] = 1.f; cl.fy1[v] = 0.f; }
        if (col_in) {
            float g2[CV];
            VecIO<float, CV>::ld(p.gy2 + jb, g2);
            VecIO<float, CV>::ld(p.gy3 + jb, cl.gy3);
            VecIO<float, CV>::ld(p.fy1 + jb, cl.fy1);
            VecIO<float, CV>::ld(p.fy2 + jb, cl.fy2);
            VecIO<float, CV>::ld(p.fy3 + jb, cl.fy3);
#pragma unroll
            for (int v = 0; v < CV; ++v) cl.gy2h[v] = g2[v] * 0.5f;        // (gx2 * gy2) * 0.5 with gx2 = 1
        }
        cl.fix_d = jb == 0;
        cl.fix_h = jb + CV == p.ny;
    }
    const int r_begin = i0 - T, r_end = i1 + T;      // rows fed to stage 0: [r_begin, r_end)
    const int n_trip = (r_end - r_begin + NS - 1) / NS;   // trips of NS rows (the last one may run past r_end: never stored)
    const int leader = lane == 0;

    // shared memory of the group: staging ring, queues, barriers
    const unsigned bars = smem_u32(gsm + Shape::OFF_BAR);
    // barriers: staging slot s -> bars + 8 s; queue q slot d: full -> bars + 8 (NSTAGE + (q NS + d) 2), empty -> + 8
    const unsigned char *const in_data = FIRST ? gsm + lane * CLB
                                               : gsm + Shape::OFF_Q + (wg - 1) * NS * Shape::QSLOT_B + lane * CLB;
    const unsigned in_bar = bars + 8u * (NSTAGE + (wg - 1) * NS * 2);     // (queue behind; unused by the first warp)
    unsigned char *const out_data = gsm + Shape::OFF_Q + wg * NS * Shape::QSLOT_B + lane * CLB;   // (queue ahead)
    const unsigned out_bar = bars + 8u * (NSTAGE + wg * NS * 2);

    RowSet<float, CV> S[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k)
#pragma unroll
        for (int v = 0; v < CV; ++v) S[k].dz[v] = S[k].ez[v] = S[k].hx[v] = S[k].hy[v] = S[k].ihx[v] = S[k].ihy[v] = S[k].naz[v] = 0.f;

    float2 negzero;
    {
        const unsigned long long z = p.negzero2;
        negzero = make_float2(__uint_as_float((unsigned)z), __uint_as_float((unsigned)(z >> 32)));
    }
    const unsigned stage0 = smem_u32(gsm);
    const int row0 = r_begin - p.row_base;           // array row of the first box
    const int n_box = (n_trip * NS + R - 1) / R;     // boxes of R rows that cover every row the trips take
    if (FIRST) {                                     // prime the staging ring
#pragma unroll
        for (int b = 0; b < NSTAGE; ++b)
            if (b < n_box) tma_issue_boxes(stage0 + b * Shape::STAGE_B, Shape::BOX_B, maps, bars + 8u * b, c0, row0 + b * R);
    }
    // run-time walk of the staging ring (first warp, !SRING): slot / row in the box / box id / parity of the slot's barrier
    int rs_slot = 0, rs_row = 0, rs_box = 0;
    unsigned rs_par = 0;
    // last warp: element offset of the row stored next (its x-th input is row r_begin + x - (G-1)K, released K rows later)
    long long off_s = (long long)(r_begin - T - p.row_base) * p.ny + jb;
    int ro = r_begin - T;
    auto st2 = [&](float *dst, const float (&d)[CV]) {      // register pairs leave as two 64-bit halves (no quad assembly)
        *reinterpret_cast<float2 *>(dst) = make_float2(d[0], d[1]);
        *reinterpret_cast<float2 *>(dst + 2) = make_float2(d[2], d[3]);
    };

#pragma unroll 1
    for (int trip = 0; trip < n_trip; ++trip) {
        const unsigned par = (unsigned)trip & 1u;    // queue barriers complete one phase per trip
        const int row_in = r_begin + trip * NS - wg * K;     // global row of this warp's input at u = 0 (row variant)
        const int sslot = trip & 1;                  // staging slot of this trip's box; its barrier's parity is (trip >> 1) & 1
        const unsigned char *const src_base = (FIRST && SRING) ? in_data + sslot * Shape::STAGE_B : in_data;
        if (FIRST && SRING) mbar_wait(bars + 8u * sslot, ((unsigned)trip >> 1) & 1u);
#pragma unroll
        for (int u = 0; u < NS; ++u) {
            // ---- take the row into register set u
            {
                constexpr int ASTRIDE = FIRST ? Shape::BOX_B : CROWB;
                const unsigned char *src = src_base + u * (FIRST ? CROWB : Shape::QSLOT_B);
                if (FIRST && !SRING) {
                    if (rs_row == 0) mbar_wait(bars + 8u * rs_slot, rs_par);
                    src = in_data + rs_slot * Shape::STAGE_B + rs_row * CROWB;
                }
                if (!FIRST) mbar_wait(in_bar + 16u * u, par);
                lds_vec<float, CV>(src + 0 * ASTRIDE, S[u].dz);
                lds_vec<float, CV>(src + 1 * ASTRIDE, S[u].hx);
                lds_vec<float, CV>(src + 2 * ASTRIDE, S[u].hy);
                lds_vec<float, CV>(src + 3 * ASTRIDE, S[u].ihx);
                lds_vec<float, CV>(src + 4 * ASTRIDE, S[u].ihy);
                lds_vec<float, CV>(src + 5 * ASTRIDE, S[u].naz);
            }
            // ---- this warp's K stages: stage s has the (x-s)-th input arriving and holds the (x-s-1)-th
#pragma unroll
            for (int s = 0; s < K; ++s) {
                const int sa = (u - s + 2 * NS) % NS, sh = (u - s - 1 + 2 * NS) % NS;
                if (COL) march_stage_pk_col(S[sa], S[sh], cl, negzero);
                else if (ROW) march_stage_pk_row(p, S[sa], S[sh], row_in + u - s, negzero);      // stage s: this warp's (x-s)-th input arrives
                else march_stage_pk<CV, false, false>(S[sa], S[sh], negzero, nullptr, nullptr);
            }
            // ---- give the input slot back: every value read from it has been used by the stages above (the warp-level
            // barrier orders the other lanes' reads before the leader's arrive)
            if (!FIRST) {
                __syncwarp();
                mbar_arrive(in_bar + 16u * u + 8u, leader);
            }
            if (FIRST && !SRING) {
                if (++rs_row == R) {                 // the box is consumed: refill its slot with the box NSTAGE ahead
                    rs_row = 0;
                    __syncwarp();
                    if (rs_box + NSTAGE < n_box)
                        tma_issue_boxes(stage0 + rs_slot * Shape::STAGE_B, Shape::BOX_B, maps, bars + 8u * rs_slot, c0,
                                        row0 + (rs_box + NSTAGE) * R);
                    ++rs_box;
                    if (++rs_slot == NSTAGE) { rs_slot = 0; rs_par ^= 1u; }
                }
            }
            // ---- the row leaving the last stage: register set (u+1) % NS
            const RowSet<float, CV> &O = S[(u + 1) % NS];
            if (LAST) {
                if (ro >= i0 && ro < i1 && col_store) {
                    st2(p.out_dz + off_s, O.dz);
                    st2(p.out_hx + off_s, O.hx);
                    st2(p.out_hy + off_s, O.hy);
                    st2(p.out_ihx + off_s, O.ihx);
                    st2(p.out_ihy + off_s, O.ihy);
                    if (p.write_ez) st2(p.out_ez + off_s, O.ez);
                }
                off_s += p.ny;
                ++ro;
            } else {
                mbar_wait(out_bar + 16u * u + 8u, par ^ 1u);           // the consumer has released the slot's previous row
                unsigned char *const dst = out_data + u * Shape::QSLOT_B;
                sts22<Shape::STS_PTX>(dst + 0 * CROWB, O.dz);
                sts22<Shape::STS_PTX>(dst + 1 * CROWB, O.hx);
                sts22<Shape::STS_PTX>(dst + 2 * CROWB, O.hy);
                sts22<Shape::STS_PTX>(dst + 3 * CROWB, O.ihx);
                sts22<Shape::STS_PTX>(dst + 4 * CROWB, O.ihy);
                sts22<Shape::STS_PTX>(dst + 5 * CROWB, O.naz);
                __syncwarp();
                mbar_arrive(out_bar + 16u * u, leader);
            }
        }
        if (FIRST && SRING) {     // the box of this trip is consumed: refill its slot with the box two trips ahead
            __syncwarp();
            if (trip + 2 < n_trip)
                tma_issue_boxes(stage0 + sslot * Shape::STAGE_B, Shape::BOX_B, maps, bars + 8u * sslot, c0, row0 + (trip + 2) * NS);
        }
    }
}

// (special strip, ordinary chunk) item w of the column variant
__device__ __forceinline__ bool decode_col_item(const MarchParams<float> &p, const int w, int &strip, int &i0, int &i1) {
    const int ncf = p.nchunks - p.n_schunks;
    if (w >= p.n_sstrips * ncf) return false;
    strip = p.sstrips[w % p.n_sstrips];
    chunk_span(p, kth_not_in(w / p.n_sstrips, p.schunks, p.n_schunks), i0, i1);
    return true;
}

// (ordinary strip, special chunk) item w of the row variant
__device__ __forceinline__ bool decode_row_item(const MarchParams<float> &p, const int w, int &strip, int &i0, int &i1) {
    const int nsf = p.nstrips - p.n_sstrips;
    if (w >= nsf * p.n_schunks) return false;
    strip = kth_not_in(w % nsf, p.sstrips, p.n_sstrips);
    chunk_span(p, p.schunks[w / nsf], i0, i1);
    return true;
}

template <typename Shape, int KIND>
__global__ void __launch_bounds__(Shape::THREADS, 1)
k_march_chain(const __grid_constant__ MarchParams<float> p, const __grid_constant__ ChainMaps maps) {
    extern __shared__ __align__(1024) unsigned char chain_smem[];
    const int lane = threadIdx.x & 31;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);      // (tells the compiler the value is warp-uniform)
    const int grp = warp / Shape::G, wg = warp % Shape::G;
    unsigned char *const gsm = chain_smem + (size_t)grp * Shape::GROUP_SMEM;
    if (wg == 0 && lane == 0) {
        const unsigned bars = smem_u32(gsm + Shape::OFF_BAR);
        for (int b = 0; b < Shape::NBAR; ++b) mbar_init(bars + 8u * b, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int strip, i0, i1;
    const int item = blockIdx.x * Shape::GROUPS + grp;
    const bool mine = KIND == K_COL ? decode_col_item(p, item, strip, i0, i1)
                    : KIND == K_ROW ? decode_row_item(p, item, strip, i0, i1)
                                    : decode_item<true>(p, item, 0, CV, Shape::T, false, strip, i0, i1);
    if (!mine) return;                               // the whole group
    if (wg == 0) chain_body<Shape, true, false, KIND>(p, maps, strip, i0, i1, lane, wg, gsm);
    else if (wg == Shape::G - 1) chain_body<Shape, false, true, KIND>(p, maps, strip, i0, i1, lane, wg, gsm);
    else chain_body<Shape, false, false, KIND>(p, maps, strip, i0, i1, lane, wg, gsm);
}

// ---- host: tensor maps
typedef CUresult (*EncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiled encoder() {
    static EncodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult res;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &res) == cudaSuccess && res == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiled>(sym);
    });
    return fn;
}

int make_map(CUtensorMap *m, const float *base, int ny, int rows, int box_rows) {
    EncodeTiled enc = encoder();
    if (enc == nullptr) { fdtd::set_error("cuTensorMapEncodeTiled is not available from this driver"); return FDTD_EUNSUPPORTED; }
    const cuuint64_t gdim[2] = {(cuuint64_t)ny, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)ny * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)(32 * CV), (cuuint32_t)box_rows};
    const cuuint32_t estride[2] = {1, 1};
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), gdim, gstride, box, estride,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { fdtd::set_error("cuTensorMapEncodeTiled failed (%d) for a %d x %d float array", (int)r, rows, ny); return FDTD_ECUDA; }
    return FDTD_OK;
}

template <typename Shape, int KIND = K_INTERIOR>
int launch_chain(const MarchParams<float> &mp, int items, cudaStream_t st) {
    if (items <= 0) return FDTD_OK;
    ChainMaps maps;
    const int rows = mp.in_hi - mp.row_base;             // array rows the pass may read
    const float *arr[NARR] = {mp.in_dz, mp.in_hx, mp.in_hy, mp.in_ihx, mp.in_ihy, mp.naz};
    for (int a = 0; a < NARR; ++a) {
        const int rc = make_map(&maps.m[a], arr[a], mp.ny, rows, Shape::R);
        if (rc != FDTD_OK) return rc;
    }
    static bool configured[64] = {false};
    static std::mutex guard;
    {
        std::lock_guard<std::mutex> lock(guard);
        int dev = 0;
        FDTD_CUDA(cudaGetDevice(&dev));
        bool &done = configured[dev >= 0 && dev < 64 ? dev : 0];
        if (!done || dev >= 64) {
            FDTD_CUDA(cudaFuncSetAttribute(k_march_chain<Shape, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, Shape::SMEM));
            done = true;
        }
    }
    const int grid = (items + Shape::GROUPS - 1) / Shape::GROUPS;
    k_march_chain<Shape, KIND><<<grid, Shape::THREADS, Shape::SMEM, st>>>(mp, maps);
    FDTD_LAUNCH_CHECK("k_march_chain");
    return FDTD_OK;
}

//                        G  K  GROUPS [R NSTAGE STS_PTX]      measured at 32768^2 (profiles/r2_chain_shapes.txt)
using Chain8 = ChainShape<2, 4, 4>;                    // depth 8, shipped: two warps of four stages, eight warps per SM -- 889 Gcell/s
using Chain8b = ChainShape<2, 4, 6, 2, 3, true>;       // twelve warps per SM: three boxes of two rows, run-time staging ring -- 885
using Chain8c = ChainShape<2, 4, 6, 3, 2, true>;       // twelve warps per SM: two boxes of three rows -- 814
using Chain8d = ChainShape<4, 2, 4>;                   // four warps of two stages, sixteen warps per SM -- 711 (twice the hand-offs per stage)
using Chain12 = ChainShape<4, 3, 3>;                   // depth 12, shipped: four warps of three stages, twelve warps per SM -- 765
using Chain12b = ChainShape<4, 3, 3, 2, 3>;            // run-time staging ring, boxes of two rows -- 718
using Chain12c = ChainShape<3, 4, 3>;                  // three warps of four stages, nine warps per SM -- 524

}  // namespace

namespace fdtd_march {

bool chain_supported(int T, bool lossy) {
    return (T == 8 || T == 12) && !lossy && encoder() != nullptr;
}

int launch_march_chain(const MarchParams<float> &mp, int T, int shape, int items, cudaStream_t st, int kind) {
    if (mp.ny % CV != 0 || (reinterpret_cast<uintptr_t>(mp.in_dz) & 15u) != 0) {
        fdtd::set_error("warp-chain pass: ny must be a multiple of 4 and the arrays 16-byte aligned");
        return FDTD_EINVAL;
    }
    if (kind == K_COL) {         // (special strip, ordinary chunk) items: the shipped shape with the column-coefficient stage
        if (T == 8) return launch_chain<Chain8, K_COL>(mp, items, st);
        if (T == 12) return launch_chain<Chain12, K_COL>(mp, items, st);
    }
    if (kind == K_ROW) {         // (ordinary strip, special chunk) items: ... with the row-coefficient stage
        if (T == 8) return launch_chain<Chain8, K_ROW>(mp, items, st);
        if (T == 12) return launch_chain<Chain12, K_ROW>(mp, items, st);
    }
    if (T == 8) {
        if (shape == 1) return launch_chain<Chain8b>(mp, items, st);
        if (shape == 2) return launch_chain<Chain8c>(mp, items, st);
        if (shape == 3) return launch_chain<Chain8d>(mp, items, st);
        return launch_chain<Chain8>(mp, items, st);
    }
    if (T == 12) {
        if (shape == 1) return launch_chain<Chain12b>(mp, items, st);
        if (shape == 2) return launch_chain<Chain12c>(mp, items, st);
        return launch_chain<Chain12>(mp, items, st);
    }
    fdtd::set_error("no warp-chain pass of depth %d", T);
    return FDTD_EUNSUPPORTED;
}

void preload_chain() {
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, k_march_chain<Chain8, K_INTERIOR>);
    cudaFuncGetAttributes(&a, k_march_chain<Chain12, K_INTERIOR>);
    cudaFuncGetAttributes(&a, k_march_chain<Chain8, K_COL>);
    cudaFuncGetAttributes(&a, k_march_chain<Chain12, K_COL>);
    cudaFuncGetAttributes(&a, k_march_chain<Chain8, K_ROW>);
    cudaFuncGetAttributes(&a, k_march_chain<Chain12, K_ROW>);
}

}  // namespace fdtd_march
