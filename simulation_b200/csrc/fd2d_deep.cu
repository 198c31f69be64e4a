// Deep passes of the fused 2D TM kernel: T = 8 or 12 full time steps per trip through HBM (float, 4 columns per lane).
//
// The register pipeline of fd2d_march.cu keeps 7 arrays x 4 columns per row in registers and T+1 rows in flight, which
// fills the register file at T = 6.  Of those seven arrays only three are on the dependent chain of a stage: dz, hx and
// hy.  ihx and ihy are ACCUMULATORS (read, add, write back; in the interior nothing else ever reads them except for the
// sign of a zero), naz is read-only.  Here they live in shared memory -- each lane owns 16 bytes per array per row and
// touches nothing else, so there is still no barrier and no bank conflict -- and the registers hold 3 arrays per row:
// 13 row sets x 12 registers at T = 12.  Ez is not stored anywhere between the two stages that use it: the held row's
// Ez is the same product naz*dz evaluated again (same operands, same bits).
//
// Why not ALL of the state in shared memory (CTA-wide TMA tiles): a stage reads and writes every array of two rows, ~48
// bytes of shared-memory traffic per cell-update; at 128 B/clk/SM that alone caps the chip at 148 x 1.965 GHz x 128 / 48
// = 775 Gcell/s -- the figure the register pipeline already reaches.  With the three chain arrays in registers the
// shared-memory traffic is 20 B per cell-update (ihx, ihy read + write, naz read twice).
//
// Per pass the real HBM traffic is still ~44 B per cell (every array once in, once out), but it now buys 12 steps
// instead of 6.  Strips overlap by 12 columns per side (104 of 128 produced), chunks are 256 rows (+12 per side).
//
// The careful kernel of these passes (k_careful2) keeps WHOLE row sets in a shared-memory ring and runs the stages in a
// rolled loop with a run-time depth: compact code, no register shifting, no spills, any T <= TMAX.  It calls the same
// march_stage<.., FAST = false> as the register-pipeline careful kernel, so the arithmetic is shared line for line.
#include "fd2d_march.cuh"

#include <mutex>

namespace {

using namespace fdtd_march;

constexpr int DV = 4;                 // columns per lane
constexpr int DLB = DV * 4;           // bytes per lane per array row
constexpr int DROWB = 32 * DLB;       // bytes per array row of a warp

// Compile-time shape of one instantiation.  Staged per row (cp.async ring of DRING rows): dz hx hy ihx ihy naz [iz nbz].
// Resident per register row set (T+1 rows, compile-time slots): naz, and whichever of ihx / ihy / ez the registers do
// not hold (KEEP bits) [lossy: iz nbz ez as well].  Shared-memory bandwidth is what bounds these kernels (128 B per
// clock per SM: one LDS.128 / STS.128 of a warp is four cycles of the data pipe), so every array kept in registers
// buys back eight (accumulators: read + write) or four (ez: the second read of naz) of the ~31 cycles a stage costs.
constexpr int KEEP_IHX = 1, KEEP_EZ = 2, KEEP_IHY = 4;

template <int T, bool LOSSY, int DRING, int WARPS_ = MAX_WARPS, int KEEP = 0>
struct DeepShape {
    static constexpr int NS = T + 1;
    static constexpr int NSTG = LOSSY ? 8 : 6;
    static constexpr bool IHX_REG = (KEEP & KEEP_IHX) != 0, IHY_REG = (KEEP & KEEP_IHY) != 0;
    static constexpr bool EZ_REG = (KEEP & KEEP_EZ) != 0 && !LOSSY;
    // resident array slots, in order: naz [ihx] [ihy] [iz nbz ez]
    static constexpr int R_NAZ = 0;
    static constexpr int R_IHX = 1;                               // (meaningful only when !IHX_REG)
    static constexpr int R_IHY = R_IHX + (IHX_REG ? 0 : 1);       // (meaningful only when !IHY_REG)
    static constexpr int R_IZ = R_IHY + (IHY_REG ? 0 : 1);
    static constexpr int R_NBZ = R_IZ + 1, R_EZ = R_IZ + 2;
    static constexpr int NRES = R_IZ + (LOSSY ? 3 : 0);
    static constexpr int SLOT = NSTG * DROWB;             // staging bytes per row
    static constexpr int RA = NS * DROWB;                 // bytes of one resident array
    static constexpr int WARP_SMEM = DRING * SLOT + NRES * RA;
    static constexpr int MAXW = (227 * 1024) / WARP_SMEM;
    static constexpr int WARPS = MAXW < WARPS_ ? MAXW : WARPS_;
};

struct DeepRow { float dz[DV], hx[DV], hy[DV], ihx[DV], ihy[DV], ez[DV]; };   // (ihx / ihy / ez: only where KEEP says so)

__device__ __forceinline__ void sts4(void *dst, const float (&d)[DV]) {
    *reinterpret_cast<float4 *>(dst) = make_float4(d[0], d[1], d[2], d[3]);
}
// values computed in register PAIRS leave as two 64-bit halves: assembling an aligned quad for a 128-bit store costs
// four moves (same bytes, same banks)
__device__ __forceinline__ void sts22(void *dst, const float (&d)[DV]) {
    *reinterpret_cast<float2 *>(dst) = make_float2(d[0], d[1]);
    *reinterpret_cast<float2 *>(reinterpret_cast<unsigned char *>(dst) + 8) = make_float2(d[2], d[3]);
}

// One stage in packed arithmetic: D, E of the arriving row A and H of the held row Hd, operation for operation what
// march_stage_pk<4, ...> does -- with the accumulators that live in shared memory read from and written back to the
// held row's resident slot.  resA / resH: this lane's 16 bytes in the resident slot of the arriving / the held row
// (array k at + k * RA).
template <typename Shape>
__device__ __forceinline__ void deep_stage(DeepRow &A, DeepRow &Hd, unsigned char *const resA, unsigned char *const resH,
                                           const float2 negzero) {
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int RA = Shape::RA;
    constexpr bool LOSSY = Shape::NSTG == 8;
    const float2 half2 = make_float2(0.5f, 0.5f), zero2 = make_float2(0.f, 0.f);
    // ---- D of the arriving row: dz = dz + 0.5*(((hy - hy[i-1]) - hx) + hx[j-1])
    const float hx_left = __shfl_up_sync(FULL, A.hx[DV - 1], 1);
#pragma unroll
    for (int v = 0; v < DV; v += 2) {
        const float2 a1 = pk_sub(make_float2(A.hy[v], A.hy[v + 1]), make_float2(Hd.hy[v], Hd.hy[v + 1]));
        const float2 a2 = pk_sub(a1, make_float2(A.hx[v], A.hx[v + 1]));
        const float2 curl = make_float2(a2.x + (v == 0 ? hx_left : A.hx[v == 0 ? 0 : v - 1]), a2.y + A.hx[v]);   // shifted pair: scalar
        const float2 dn = pk_add(make_float2(A.dz[v], A.dz[v + 1]), pk_mul(half2, curl, negzero));
        A.dz[v] = dn.x; A.dz[v + 1] = dn.y;
    }
    // ---- E of both rows
    float ezA[DV], ezH[DV], nzA[DV];
    lds_vec<float, DV>(resA + Shape::R_NAZ * RA, nzA);
    if constexpr (LOSSY) {          // ez = naz*(dz - iz); iz = iz + nbz*ez -- ez kept for the next trip, iz in place
        float iz[DV], nb[DV];
        lds_vec<float, DV>(resA + Shape::R_IZ * RA, iz);
        lds_vec<float, DV>(resA + Shape::R_NBZ * RA, nb);
#pragma unroll
        for (int v = 0; v < DV; v += 2) {
            const float2 i0 = make_float2(iz[v], iz[v + 1]);
            const float2 a = pk_mul(make_float2(nzA[v], nzA[v + 1]), pk_sub(make_float2(A.dz[v], A.dz[v + 1]), i0), negzero);
            const float2 i2 = pk_add(i0, pk_mul(make_float2(nb[v], nb[v + 1]), a, negzero));
            iz[v] = i2.x; iz[v + 1] = i2.y;
            ezA[v] = a.x; ezA[v + 1] = a.y;
        }
        sts22(resA + Shape::R_IZ * RA, iz);
        sts22(resA + Shape::R_EZ * RA, ezA);
        lds_vec<float, DV>(resH + Shape::R_EZ * RA, ezH);
    } else if constexpr (Shape::EZ_REG) {     // Ez travels with the row set
#pragma unroll
        for (int v = 0; v < DV; v += 2) {
            const float2 a = pk_mul(make_float2(nzA[v], nzA[v + 1]), make_float2(A.dz[v], A.dz[v + 1]), negzero);
            A.ez[v] = a.x; A.ez[v + 1] = a.y;
            ezA[v] = a.x; ezA[v + 1] = a.y; ezH[v] = Hd.ez[v]; ezH[v + 1] = Hd.ez[v + 1];
        }
    } else {                                  // the held row's Ez is naz*dz evaluated again (same operands, same bits)
        float nzH[DV];
        lds_vec<float, DV>(resH + Shape::R_NAZ * RA, nzH);
#pragma unroll
        for (int v = 0; v < DV; v += 2) {
            const float2 a = pk_mul(make_float2(nzA[v], nzA[v + 1]), make_float2(A.dz[v], A.dz[v + 1]), negzero);
            const float2 h = pk_mul(make_float2(nzH[v], nzH[v + 1]), make_float2(Hd.dz[v], Hd.dz[v + 1]), negzero);
            ezA[v] = a.x; ezA[v + 1] = a.y; ezH[v] = h.x; ezH[v + 1] = h.y;
        }
    }
    // ---- H of the held row: ihx += cm; ihy += cn; hx = hx + (0.5*cm + 0*ihx); hy = hy - (0.5*cn + 0*ihy)
    const float ez_right = __shfl_down_sync(FULL, ezH[0], 1);
    float ihx[DV], ihy[DV];
    if constexpr (Shape::IHX_REG) {
#pragma unroll
        for (int v = 0; v < DV; ++v) ihx[v] = Hd.ihx[v];
    } else {
        lds_vec<float, DV>(resH + Shape::R_IHX * RA, ihx);
    }
    if constexpr (Shape::IHY_REG) {
#pragma unroll
        for (int v = 0; v < DV; ++v) ihy[v] = Hd.ihy[v];
    } else {
        lds_vec<float, DV>(resH + Shape::R_IHY * RA, ihy);
    }
#pragma unroll
    for (int v = 0; v < DV; v += 2) {
        const float2 e = make_float2(ezH[v], ezH[v + 1]);
        const float2 cm = make_float2(ezH[v] - ezH[v + 1], ezH[v + 1] - (v + 2 < DV ? ezH[v + 2 < DV ? v + 2 : v] : ez_right));   // shifted pair: scalar
        const float2 cn = pk_sub(e, make_float2(ezA[v], ezA[v + 1]));
        const float2 sx = pk_add(make_float2(ihx[v], ihx[v + 1]), cm);
        const float2 sy = pk_add(make_float2(ihy[v], ihy[v + 1]), cn);
        const float2 tx = pk_add(pk_mul(half2, cm, negzero), pk_mul(zero2, sx, negzero));
        const float2 ty = pk_add(pk_mul(half2, cn, negzero), pk_mul(zero2, sy, negzero));
        const float2 hx2 = pk_add(make_float2(Hd.hx[v], Hd.hx[v + 1]), tx);
        const float2 hy2 = pk_sub(make_float2(Hd.hy[v], Hd.hy[v + 1]), ty);
        ihx[v] = sx.x; ihx[v + 1] = sx.y; ihy[v] = sy.x; ihy[v + 1] = sy.y;
        Hd.hx[v] = hx2.x; Hd.hx[v + 1] = hx2.y; Hd.hy[v] = hy2.x; Hd.hy[v + 1] = hy2.y;
    }
    if constexpr (Shape::IHX_REG) {
#pragma unroll
        for (int v = 0; v < DV; ++v) Hd.ihx[v] = ihx[v];
    } else {
        sts22(resH + Shape::R_IHX * RA, ihx);
    }
    if constexpr (Shape::IHY_REG) {
#pragma unroll
        for (int v = 0; v < DV; ++v) Hd.ihy[v] = ihy[v];
    } else {
        sts22(resH + Shape::R_IHY * RA, ihy);
    }
}

// The march of one interior warp over its (strip, chunk): every column (halo included) is an ordinary cell and every
// row touched is an ordinary stored row (the host's classification guarantees it), so there are no masks at all.
template <typename Shape, int T, bool LOSSY, int DRING>
__device__ __forceinline__ void deep_body(const MarchParams<float> &p, const int strip, const int i0, const int i1,
                                          const int lane, unsigned char *const smem) {
    constexpr int W = 32 * DV;
    constexpr int HALO = ((T + DV - 1) / DV) * DV;
    constexpr int USE = W - 2 * HALO;
    constexpr int NS = Shape::NS, SLOT = Shape::SLOT, RA = Shape::RA;

    const int c0 = strip * USE - HALO;               // first column of the strip (halo included)
    const int jb = c0 + lane * DV;                   // first column of this lane
    const bool col_store = (lane * DV >= HALO) && (lane * DV + DV <= W - HALO);

    DeepRow S[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k)
#pragma unroll
        for (int v = 0; v < DV; ++v) S[k].dz[v] = S[k].hx[v] = S[k].hy[v] = S[k].ihx[v] = S[k].ihy[v] = S[k].ez[v] = 0.f;

    unsigned char *const lane_ring = smem + lane * DLB;                     // staging ring, this lane's column
    unsigned char *const lane_res = smem + DRING * SLOT + lane * DLB;       // resident arrays, this lane's column
    // rows above the chunk (pipeline warm-up) read resident slots no row has been taken into yet
#pragma unroll
    for (int k = 0; k < Shape::NRES * NS; ++k) *reinterpret_cast<float4 *>(lane_res + k * DROWB) = make_float4(0.f, 0.f, 0.f, 0.f);

    const int r_begin = i0 - T, r_end = i1 + T;      // rows fed to stage 0: [r_begin, r_end)
    long long off_f = (long long)(r_begin - p.row_base) * p.ny + jb;       // element offset of the row fetched next
    long long off_s = (long long)(r_begin - T - p.row_base) * p.ny + jb;   // ... of the row stored next
    auto fetch = [&](const int k) {
        unsigned char *dst = lane_ring + k * SLOT;
        cp_async<DLB>(dst + 0 * DROWB, p.in_dz + off_f, DLB);
        cp_async<DLB>(dst + 1 * DROWB, p.in_hx + off_f, DLB);
        cp_async<DLB>(dst + 2 * DROWB, p.in_hy + off_f, DLB);
        cp_async<DLB>(dst + 3 * DROWB, p.in_ihx + off_f, DLB);
        cp_async<DLB>(dst + 4 * DROWB, p.in_ihy + off_f, DLB);
        cp_async<DLB>(dst + 5 * DROWB, p.naz + off_f, DLB);
        if (LOSSY) {
            cp_async<DLB>(dst + 6 * DROWB, p.in_iz + off_f, DLB);
            cp_async<DLB>(dst + 7 * DROWB, p.nbz + off_f, DLB);
        }
        cp_async_commit();
        off_f += p.ny;
    };
    // a landed row: the chain arrays into the register set, the rest into its resident slot
    auto take = [&](const int k, DeepRow &row, const int set) {
        const unsigned char *src = lane_ring + k * SLOT;
        lds_vec<float, DV>(src + 0 * DROWB, row.dz);
        lds_vec<float, DV>(src + 1 * DROWB, row.hx);
        lds_vec<float, DV>(src + 2 * DROWB, row.hy);
        unsigned char *res = lane_res + set * DROWB;
        float t[DV];
        if constexpr (Shape::IHX_REG) lds_vec<float, DV>(src + 3 * DROWB, row.ihx);
        else { lds_vec<float, DV>(src + 3 * DROWB, t); sts4(res + Shape::R_IHX * RA, t); }
        if constexpr (Shape::IHY_REG) lds_vec<float, DV>(src + 4 * DROWB, row.ihy);
        else { lds_vec<float, DV>(src + 4 * DROWB, t); sts4(res + Shape::R_IHY * RA, t); }
        lds_vec<float, DV>(src + 5 * DROWB, t); sts4(res + Shape::R_NAZ * RA, t);
        if (LOSSY) {
            lds_vec<float, DV>(src + 6 * DROWB, t); sts4(res + Shape::R_IZ * RA, t);
            lds_vec<float, DV>(src + 7 * DROWB, t); sts4(res + Shape::R_NBZ * RA, t);
        }
    };
    float2 negzero;
    {
        const unsigned long long z = p.negzero2;
        negzero = make_float2(__uint_as_float((unsigned)z), __uint_as_float((unsigned)(z >> 32)));
    }
    auto st2 = [&](float *dst, const float (&d)[DV]) {      // register pairs leave as two 64-bit halves (no quad assembly)
        *reinterpret_cast<float2 *>(dst) = make_float2(d[0], d[1]);
        *reinterpret_cast<float2 *>(dst + 2) = make_float2(d[2], d[3]);
    };
    auto store_row = [&](const DeepRow &O, const int ro, const int set) {
        if (ro >= i0 && ro < i1 && col_store) {
            const unsigned char *res = lane_res + set * DROWB;
            st2(p.out_dz + off_s, O.dz);
            st2(p.out_hx + off_s, O.hx);
            st2(p.out_hy + off_s, O.hy);
            float t[DV];
            if constexpr (Shape::IHX_REG) st2(p.out_ihx + off_s, O.ihx);
            else { lds_vec<float, DV>(res + Shape::R_IHX * RA, t); VecIO<float, DV>::st(p.out_ihx + off_s, t); }
            if constexpr (Shape::IHY_REG) st2(p.out_ihy + off_s, O.ihy);
            else { lds_vec<float, DV>(res + Shape::R_IHY * RA, t); VecIO<float, DV>::st(p.out_ihy + off_s, t); }
            if (LOSSY) {
                lds_vec<float, DV>(res + Shape::R_IZ * RA, t); VecIO<float, DV>::st(p.out_iz + off_s, t);
                if (p.write_ez) { lds_vec<float, DV>(res + Shape::R_EZ * RA, t); VecIO<float, DV>::st(p.out_ez + off_s, t); }
            } else if (p.write_ez) {
                if constexpr (Shape::EZ_REG) {
                    st2(p.out_ez + off_s, O.ez);
                } else {                          // the last pass evaluates naz*dz once more (same operands, same bits)
                    float e[DV];
                    lds_vec<float, DV>(res + Shape::R_NAZ * RA, t);
#pragma unroll
                    for (int v = 0; v < DV; v += 2) {
                        const float2 a = pk_mul(make_float2(t[v], t[v + 1]), make_float2(O.dz[v], O.dz[v + 1]), negzero);
                        e[v] = a.x; e[v + 1] = a.y;
                    }
                    st2(p.out_ez + off_s, e);
                }
            }
        }
        off_s += p.ny;
    };

    // Staging ring.  DRING >= 3: DRING-1 rows in flight, the slot consumed one sub-iteration ago is refilled right after the
    // take.  DRING == 2 (deepest shapes, where shared memory decides how many warps fit): the slot just consumed is
    // refilled at the END of the sub-iteration -- every value read from it has been used by then -- so one row is in
    // flight during the T stages and two across the boundary.
    constexpr bool LATE = DRING == 2;
    constexpr int AHEAD = LATE ? 2 : DRING - 1;       // rows fetched before the first take
#pragma unroll
    for (int k = 0; k < AHEAD; ++k) fetch(k);
    int slot = 0;                                     // staging slot of the row consumed next
#pragma unroll 1                                      // the body is T*(T+1) stages already
    for (int r = r_begin; r < r_end; r += NS) {
#pragma unroll
        for (int u = 0; u < NS; ++u) {
            const int rr = r + u;                     // global row arriving at stage 0 (may overrun r_end)
            cp_async_wait<AHEAD - 1>();               // the oldest pending row has landed
            take(slot, S[u], u);
            if (!LATE) fetch(slot == 0 ? DRING - 1 : slot - 1);  // refill the slot consumed one sub-iteration ago with row rr + DRING - 1
#pragma unroll
            for (int s = 0; s < T; ++s) {             // stage s: row rr-s arrives, row rr-s-1 is held
                const int sa = (u - s + 2 * NS) % NS, sh = (u - s - 1 + 2 * NS) % NS;
                deep_stage<Shape>(S[sa], S[sh], lane_res + sa * DROWB, lane_res + sh * DROWB, negzero);
            }
            store_row(S[(u + 1) % NS], rr - T, (u + 1) % NS);     // the set held by the last stage: row rr-T at time t+T
            if (LATE) fetch(slot);                    // row rr + 2 into the slot taken above
            slot = (slot + 1 == DRING) ? 0 : slot + 1;
        }
    }
    cp_async_wait<0>();
}

// Register budget stated directly.  Each of the four sub-partitions of an SM has its own file of 16384 registers and
// the warps of a CTA are dealt round-robin, so W warps per CTA put ceil(W/4) on one file: that, not 65536 / W, bounds the
// registers per thread (9..12 warps: 168).
constexpr int deep_maxnreg(int warps) {
    const int r = (16384 / ((warps + 3) / 4) / 32) / 8 * 8;
    return r > 255 ? 255 : r;
}
template <int T, bool LOSSY, int DRING, int WARPS, int KEEP>
__global__ void __maxnreg__(deep_maxnreg(DeepShape<T, LOSSY, DRING, WARPS, KEEP>::WARPS))
k_march_deep(const __grid_constant__ MarchParams<float> p) {
    using Shape = DeepShape<T, LOSSY, DRING, WARPS, KEEP>;
    extern __shared__ __align__(16) unsigned char deep_smem[];
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    unsigned char *const mine = deep_smem + (size_t)(threadIdx.x >> 5) * Shape::WARP_SMEM;
    int strip, i0, i1;
    if (!decode_item<true>(p, w, 0, DV, T, LOSSY, strip, i0, i1)) return;
    deep_body<Shape, T, LOSSY, DRING>(p, strip, i0, i1, lane, mine);
}

// ------------------------------------------------------------------------------------------------------------------
// Careful (edge-aware) warps with the whole row pipeline in shared memory.  Ring of T+3 rows per warp (T+1 in the
// pipeline + 2 in flight), 7 (lossy: 9) arrays per row, each lane its own V elements; the stages run in a rolled loop
// and load / store the two rows they touch.  Depth T is a run-time value.
enum { C_DZ = 0, C_EZ, C_HX, C_HY, C_IHX, C_IHY, C_NAZ, C_IZ, C_NBZ };

template <typename real, int V, int MODE>
struct CarefulShape {
    static constexpr bool LOSSY = (MODE & 1) != 0;
    static constexpr int NARR = LOSSY ? 9 : 7;
    static constexpr int LB = V * (int)sizeof(real), ROWB = 32 * LB;
    static int warp_smem(int T) { return (T + 3) * NARR * ROWB; }
};

template <typename real, int V>
__device__ __forceinline__ void sts_vec(void *dst, const real (&d)[V]) {
    real *q = reinterpret_cast<real *>(dst);
    if constexpr (sizeof(real) * V == 16) {
        float4 t;
        real *tq = reinterpret_cast<real *>(&t);
#pragma unroll
        for (int v = 0; v < V; ++v) tq[v] = d[v];
        *reinterpret_cast<float4 *>(dst) = t;
    } else if constexpr (sizeof(real) * V == 8) {
        float2 t;
        real *tq = reinterpret_cast<real *>(&t);
#pragma unroll
        for (int v = 0; v < V; ++v) tq[v] = d[v];
        *reinterpret_cast<float2 *>(dst) = t;
    } else {
        q[0] = d[0];
    }
}

template <typename real, int V, int MODE>
__device__ __forceinline__ void careful2_body(const MarchParams<real> &p, const int T, const int strip, const int i0,
                                              const int i1, const int lane, unsigned char *const ring) {
    using Shape = CarefulShape<real, V, MODE>;
    constexpr bool LOSSY = Shape::LOSSY;
    constexpr int W = 32 * V, NARR = Shape::NARR, LB = Shape::LB, ROWB = Shape::ROWB;
    const int HALO = ((T + V - 1) / V) * V;          // recomputed columns per side: >= T, multiple of V (aligned vectors)
    const int USE = W - 2 * HALO;                    // columns a strip produces
    const int R = T + 3;                             // ring rows

    const int c0 = strip * USE - HALO;               // first column of the strip (halo included)
    const int jb = c0 + lane * V;                    // first column of this lane
    const bool col_in = (jb >= 0) && (jb + V <= p.ny);
    const bool col_store = col_in && (lane * V >= HALO) && (lane * V + V <= W - HALO);

    ColCoef<real, V> c;
    c.dmask = c.hmask = 0;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        c.gy2[v] = c.gy3[v] = c.fy2[v] = c.fy3[v] = real(1);
        c.fy1[v] = real(0);
    }
    if (col_in) {
        VecIO<real, V>::ld(p.gy2 + jb, c.gy2);
        VecIO<real, V>::ld(p.gy3 + jb, c.gy3);
        VecIO<real, V>::ld(p.fy1 + jb, c.fy1);
        VecIO<real, V>::ld(p.fy2 + jb, c.fy2);
        VecIO<real, V>::ld(p.fy3 + jb, c.fy3);
#pragma unroll
        for (int v = 0; v < V; ++v) {
            if (jb + v >= 1) c.dmask |= 1u << v;
            if (jb + v <= p.ny - 2) c.hmask |= 1u << v;
        }
    }
    // warp-uniform "does this strip touch a special column" flags
    const int ja = p.npml - 1, jz = p.ny - p.npml;
    const bool tf_cols = p.tfsf && ((ja - 1 >= c0 && ja - 1 < c0 + W) || (ja >= c0 && ja < c0 + W) || (jz >= c0 && jz < c0 + W));
    const bool src_cols = (p.src_i >= 0) && (p.src_j >= c0 && p.src_j < c0 + W);

    unsigned char *const lane_ring = ring + lane * LB;
    auto at = [&](const int slot, const int arr) -> unsigned char * { return lane_ring + (slot * NARR + arr) * ROWB; };
    {   // rows above the chunk (pipeline warm-up) are zero rows
        real z[V];
#pragma unroll
        for (int v = 0; v < V; ++v) z[v] = real(0);
        for (int k = 0; k < R * NARR; ++k) sts_vec<real, V>(lane_ring + k * ROWB, z);
    }
    const int r_begin = i0 - T, r_end = i1 + T;      // rows fed to stage 0: [r_begin, r_end)
    long long off_f = (long long)(r_begin - p.row_base) * p.ny + jb;
    long long off_s = (long long)(r_begin - T - p.row_base) * p.ny + jb;
    int g_f = r_begin;                               // global row fetched next
    // asynchronous copies of global row g_f into ring slot k (zero fill outside the stored rows / the grid)
    auto fetch = [&](const int k) {
        const bool ok = (g_f >= p.in_lo) && (g_f < p.in_hi) && (g_f < r_end) && col_in;
        const long long off = ok ? off_f : 0;
        const int nb = ok ? LB : 0;
        cp_async<LB>(at(k, C_DZ), p.in_dz + off, nb);
        cp_async<LB>(at(k, C_HX), p.in_hx + off, nb);
        cp_async<LB>(at(k, C_HY), p.in_hy + off, nb);
        cp_async<LB>(at(k, C_IHX), p.in_ihx + off, nb);
        cp_async<LB>(at(k, C_IHY), p.in_ihy + off, nb);
        cp_async<LB>(at(k, C_NAZ), p.naz + off, nb);
        if (LOSSY) {
            cp_async<LB>(at(k, C_IZ), p.in_iz + off, nb);
            cp_async<LB>(at(k, C_NBZ), p.nbz + off, nb);
        }
        cp_async_commit();
        off_f += p.ny;
        ++g_f;
    };
    auto store_row = [&](const int slot, const int ro) {
        if (ro >= i0 && ro < i1 && col_store) {
            real dz[V], hx[V], hy[V], ihx[V], ihy[V], iz[V], ez[V];
            lds_vec<real, V>(at(slot, C_DZ), dz);   lds_vec<real, V>(at(slot, C_HX), hx);
            lds_vec<real, V>(at(slot, C_HY), hy);   lds_vec<real, V>(at(slot, C_IHX), ihx);
            lds_vec<real, V>(at(slot, C_IHY), ihy);
            VecIO<real, V>::st(p.out_dz + off_s, dz);
            if (p.write_ez) {
                lds_vec<real, V>(at(slot, C_EZ), ez);
                VecIO<real, V>::st(p.out_ez + off_s, ez);
            }
            VecIO<real, V>::st(p.out_hx + off_s, hx);
            VecIO<real, V>::st(p.out_hy + off_s, hy);
            VecIO<real, V>::st(p.out_ihx + off_s, ihx);
            VecIO<real, V>::st(p.out_ihy + off_s, ihy);
            if (LOSSY) {
                lds_vec<real, V>(at(slot, C_IZ), iz);
                VecIO<real, V>::st(p.out_iz + off_s, iz);
            } else {
#pragma unroll
                for (int v = 0; v < V; ++v) iz[v] = real(0);
            }
            // halo exchange fused into the pass: peer stores over NVLink, row by row
            if (p.push) push_row<real, V, LOSSY>(p, off_s, ro, dz, hx, hy, ihx, ihy, iz);
        }
        off_s += p.ny;
    };

    int sf = 0;                                      // slot fetched next
    fetch(sf); sf = 1;
    fetch(sf); sf = 2;
    int sr = 0;                                      // slot of the row arriving now
#pragma unroll 1
    for (int rr = r_begin; rr < r_end; ++rr) {
        cp_async_wait<1>();                          // row rr has landed (row rr+1 may still be in flight)
        fetch(sf);                                   // row rr+2 into the slot row rr-T-1 left one trip ago
        sf = (sf + 1 == R) ? 0 : sf + 1;
        int sa = sr;
#pragma unroll 1
        for (int s = 0; s < T; ++s) {
            const int sh = (sa == 0) ? R - 1 : sa - 1;
            RowSet<real, V> A, Hd;
            lds_vec<real, V>(at(sa, C_DZ), A.dz);   lds_vec<real, V>(at(sa, C_HX), A.hx);
            lds_vec<real, V>(at(sa, C_HY), A.hy);   lds_vec<real, V>(at(sa, C_NAZ), A.naz);
            if (LOSSY) { lds_vec<real, V>(at(sa, C_IZ), A.iz); lds_vec<real, V>(at(sa, C_NBZ), A.nbz); }
            lds_vec<real, V>(at(sh, C_EZ), Hd.ez);  lds_vec<real, V>(at(sh, C_HX), Hd.hx);
            lds_vec<real, V>(at(sh, C_HY), Hd.hy);  lds_vec<real, V>(at(sh, C_IHX), Hd.ihx);
            lds_vec<real, V>(at(sh, C_IHY), Hd.ihy);
            march_stage<real, V, MODE, false>(p, c, A, Hd, rr - s, s, jb, tf_cols, src_cols);
            sts_vec<real, V>(at(sa, C_DZ), A.dz);   sts_vec<real, V>(at(sa, C_EZ), A.ez);
            if (LOSSY) sts_vec<real, V>(at(sa, C_IZ), A.iz);
            sts_vec<real, V>(at(sh, C_HX), Hd.hx);  sts_vec<real, V>(at(sh, C_HY), Hd.hy);
            sts_vec<real, V>(at(sh, C_IHX), Hd.ihx); sts_vec<real, V>(at(sh, C_IHY), Hd.ihy);
            sa = sh;
        }
        store_row(sa, rr - T);                       // the row held by the last stage: row rr-T at time t+T
        sr = (sr + 1 == R) ? 0 : sr + 1;
    }
    cp_async_wait<0>();
}

constexpr int CAREFUL2_WARPS = 4;

template <typename real, int V, int MODE>
__global__ void __launch_bounds__(CAREFUL2_WARPS * 32)
k_careful2(const __grid_constant__ MarchParams<real> p, const int all_careful, const int T) {
    extern __shared__ __align__(16) unsigned char careful_smem[];
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    unsigned char *const ring = careful_smem + (size_t)(threadIdx.x >> 5) * ((T + 3) * CarefulShape<real, V, MODE>::NARR * CarefulShape<real, V, MODE>::ROWB);
    int strip, i0, i1;
    if (!decode_item<false>(p, w, all_careful, V, T, (MODE & 1) != 0, strip, i0, i1)) return;
    halo_wait(p, lane);
    careful2_body<real, V, MODE>(p, T, strip, i0, i1, lane, ring);
    halo_signal(p, lane);
}

// dynamic shared memory opt-in, once per kernel and device (shared by every host thread)
template <typename K>
int opt_in_smem(K kernel, size_t smem, size_t (&configured)[64], std::mutex &guard) {
    if (smem <= 48 * 1024) return FDTD_OK;
    std::lock_guard<std::mutex> lock(guard);
    int dev = 0;
    FDTD_CUDA(cudaGetDevice(&dev));
    size_t &opted = configured[dev >= 0 && dev < 64 ? dev : 0];
    if (smem > opted || dev >= 64) {
        FDTD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        opted = smem;
    }
    return FDTD_OK;
}

template <int V, int MODE>
int launch_careful2_k(const MarchParams<float> &mp, int T, int items, int all_careful, cudaStream_t st, int warps_per_cta) {
    if (items <= 0) return FDTD_OK;
    const size_t per_warp = (size_t)CarefulShape<float, V, MODE>::warp_smem(T);
    int warps = (warps_per_cta >= 1 && warps_per_cta <= CAREFUL2_WARPS) ? warps_per_cta : CAREFUL2_WARPS;
    while (warps > 1 && (size_t)warps * per_warp > 220 * 1024) --warps;
    const size_t smem = (size_t)warps * per_warp;
    static size_t configured[64] = {0};
    static std::mutex guard;
    const int rc = opt_in_smem(k_careful2<float, V, MODE>, smem, configured, guard);
    if (rc != FDTD_OK) return rc;
    const int grid = (items + warps - 1) / warps;
    k_careful2<float, V, MODE><<<grid, warps * 32, smem, st>>>(mp, all_careful, T);
    FDTD_LAUNCH_CHECK("k_careful2");
    return FDTD_OK;
}

template <int T, bool LOSSY, int DRING, int WARPS, int KEEP>
int launch_deep_interior(const MarchParams<float> &mp, int items, cudaStream_t st) {
    if (items <= 0) return FDTD_OK;
    using Shape = DeepShape<T, LOSSY, DRING, WARPS, KEEP>;
    int warps = (g_tune.warps >= 1 && g_tune.warps <= Shape::WARPS) ? g_tune.warps : Shape::WARPS;
    const size_t smem = (size_t)warps * Shape::WARP_SMEM;
    static size_t configured[64] = {0};
    static std::mutex guard;
    const int rc = opt_in_smem(k_march_deep<T, LOSSY, DRING, WARPS, KEEP>, smem, configured, guard);
    if (rc != FDTD_OK) return rc;
    const int grid = (items + warps - 1) / warps;
    k_march_deep<T, LOSSY, DRING, WARPS, KEEP><<<grid, warps * 32, smem, st>>>(mp);
    FDTD_LAUNCH_CHECK("k_march_deep");
    return FDTD_OK;
}

}  // namespace

namespace fdtd_march {

bool deep_supported(int T, bool lossy) { return (T == 8 || T == 12) && !lossy; }

int launch_careful2(MarchParams<float> &mp, int T, bool lossy, int items, int all_careful, cudaStream_t st, int V, int warps_per_cta) {
    if (T < 1 || T > TMAX) { fdtd::set_error("careful kernel: depth %d outside [1, %d]", T, TMAX); return FDTD_EINVAL; }
    const int w = warps_per_cta;
    if (V == 2) return lossy ? launch_careful2_k<2, 1>(mp, T, items, all_careful, st, w) : launch_careful2_k<2, 0>(mp, T, items, all_careful, st, w);
    if (V != DV) { fdtd::set_error("ring careful kernel: vector width %d (2 or 4)", V); return FDTD_EUNSUPPORTED; }
    return lossy ? launch_careful2_k<DV, 1>(mp, T, items, all_careful, st, w) : launch_careful2_k<DV, 0>(mp, T, items, all_careful, st, w);
}

int launch_march_deep(MarchParams<float> &mp, int T, bool lossy, cudaStream_t st) {
    if (!deep_supported(T, lossy)) { fdtd::set_error("no deep pass of depth %d%s", T, lossy ? " (lossy)" : ""); return FDTD_EUNSUPPORTED; }
    const bool chain = (g_tune.variant == 0 || g_tune.variant >= 10) && chain_supported(T, lossy);
    // PML-column strips x ordinary chunks at interior speed (classify_pass withdraws it where it cannot apply)
    mp.col_fast = chain && g_tune.col_fast && !mp.tfsf ? 1 : 0;
    // PML-row chunks x ordinary strips likewise; not on a slab with the fused exchange (its edge chunks read ghost rows
    // and push rows: the handshake lives in the careful kernel)
    mp.row_fast = chain && g_tune.row_fast && !mp.tfsf && !mp.halo_on ? 1 : 0;
    const PassCounts pc = classify_pass(mp, DV, T);
    if (pc.all_careful) return launch_careful2(mp, T, lossy, pc.n_careful, 1, st);
    auto launch_interior = [&]() -> int {
        // the warp-chain kernel of fd2d_chain.cu (TMA-fed pipeline of warps) is the shipped interior kernel of both
        // depths; variants 1..3 select the shared-memory-accumulator kernels of this file (also the fallback when the
        // driver has no tensor-map encoder), variants >= 10 the other chain shapes
        const int v = g_tune.variant;
        if (chain) return launch_march_chain(mp, T, v >= 10 ? v - 10 : 0, pc.n_fast, st);
        if (T == 12) return launch_deep_interior<12, false, 2, 8, 0>(mp, pc.n_fast, st);
        switch (v) {
            case 1: return launch_deep_interior<8, false, 3, 8, 0>(mp, pc.n_fast, st);
            case 2: return launch_deep_interior<8, false, 2, 8, KEEP_IHX>(mp, pc.n_fast, st);
            default: return launch_deep_interior<8, false, 3, 8, KEEP_IHX>(mp, pc.n_fast, st);
        }
    };
    // the careful kernel is small (edges only): fork it onto a side stream so the interior kernel backfills the SMs it
    // leaves idle, and join before the next pass
    SideStream *side = (pc.n_careful > 0 && pc.n_fast > 0 && g_tune.serial >= 2) ? side_stream(st) : nullptr;
    if (side != nullptr && g_tune.serial == 3) {
        // Backfill: the interior kernel goes first and takes every SM (one CTA each: 184 KB of its shared memory); the
        // careful kernel follows on a default-priority stream with ONE warp per CTA (39 KB at depth 8), so a careful
        // CTA fits into what an interior CTA leaves of an SM and its instructions ride in the interior's idle issue
        // slots -- instead of four-warp careful CTAs owning SMs outright for their latency-bound march.
        FDTD_CUDA(cudaEventRecord(side->fork, st));
        int rc = launch_interior();
        if (rc != FDTD_OK) return rc;
        FDTD_CUDA(cudaStreamWaitEvent(side->backfill, side->fork, 0));
        rc = launch_careful2(mp, T, lossy, pc.n_careful, 0, side->backfill, DV, 1);
        if (rc != FDTD_OK) return rc;
        if (pc.n_col > 0) {
            rc = launch_march_chain(mp, T, 0, pc.n_col, side->backfill, 1);
            if (rc != FDTD_OK) return rc;
        }
        if (pc.n_row > 0) {
            rc = launch_march_chain(mp, T, 0, pc.n_row, side->backfill, 2);
            if (rc != FDTD_OK) return rc;
        }
        FDTD_CUDA(cudaEventRecord(side->join, side->backfill));
        FDTD_CUDA(cudaStreamWaitEvent(st, side->join, 0));
        return FDTD_OK;
    }
    if (side == nullptr) {
        int rc = launch_careful2(mp, T, lossy, pc.n_careful, 0, st);
        if (rc != FDTD_OK) return rc;
        if (pc.n_col > 0) {
            rc = launch_march_chain(mp, T, 0, pc.n_col, st, 1);
            if (rc != FDTD_OK) return rc;
        }
        if (pc.n_row > 0) {
            rc = launch_march_chain(mp, T, 0, pc.n_row, st, 2);
            if (rc != FDTD_OK) return rc;
        }
        return launch_interior();
    }
    FDTD_CUDA(cudaEventRecord(side->fork, st));
    FDTD_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
    int rc = launch_careful2(mp, T, lossy, pc.n_careful, 0, side->stream);
    if (rc != FDTD_OK) return rc;
    if (pc.n_col > 0) {
        rc = launch_march_chain(mp, T, 0, pc.n_col, side->stream, 1);
        if (rc != FDTD_OK) return rc;
    }
    if (pc.n_row > 0) {
        rc = launch_march_chain(mp, T, 0, pc.n_row, side->stream, 2);
        if (rc != FDTD_OK) return rc;
    }
    FDTD_CUDA(cudaEventRecord(side->join, side->stream));
    rc = launch_interior();
    if (rc != FDTD_OK) return rc;
    FDTD_CUDA(cudaStreamWaitEvent(st, side->join, 0));
    return FDTD_OK;
}

void preload_deep(bool lossy) {
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, k_march_deep<12, false, 2, 8, 0>);
    cudaFuncGetAttributes(&a, k_march_deep<8, false, 3, 8, KEEP_IHX>);
    cudaFuncGetAttributes(&a, k_careful2<float, DV, 0>);
    cudaFuncGetAttributes(&a, k_careful2<float, 2, 0>);
    preload_chain();
    if (lossy) {
        cudaFuncGetAttributes(&a, k_careful2<float, DV, 1>);
        cudaFuncGetAttributes(&a, k_careful2<float, 2, 1>);
    }
}

}  // namespace fdtd_march
