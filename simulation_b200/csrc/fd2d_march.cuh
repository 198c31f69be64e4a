// Shared pieces of the fused 2D TM passes (fd2d_march.cu: register-pipeline kernels of depth <= 8; fd2d_deep.cu: deep
// passes with shared-memory-resident accumulators): the kernel parameter block, vector / cp.async helpers, the register
// row set, the careful (edge-aware) stage, the packed-arithmetic helpers and the host-side launch classification.
#pragma once
#include "common.cuh"

namespace fdtd_march {

constexpr int TMAX = 12;         // deepest pass (fd2d_deep.cu; the register-pipeline kernels of fd2d_march.cu stop at 8)
constexpr int NFMAX = 3;         // frequencies of the fused running DFT (the reference uses 3 everywhere)
constexpr int MAX_SPECIAL = 64;  // most edge / TFSF / halo strips (or chunks) a split launch can list
constexpr int MAX_PAIRS = 8;     // most single (strip, chunk) cells handed to the careful kernel (the point source)
constexpr int RING = 4;          // rows of the cp.async staging ring of the register-pipeline kernels (bounds the fetch run-ahead)
constexpr int MAX_WARPS = 8;     // warps per CTA are independent; a CTA only groups neighbouring strips for L1 locality

template <typename real>
struct MarchParams {
    const real *in_dz, *in_hx, *in_hy, *in_ihx, *in_ihy, *in_iz;
    real *out_dz, *out_ez, *out_hx, *out_hy, *out_ihx, *out_ihy, *out_iz;
    const real *naz, *nbz;
    const real *gx2, *gx3, *fx1, *fx2, *fx3;   // indexed by GLOBAL row
    const real *gy2, *gy3, *fy1, *fy2, *fy3;   // indexed by column
    int nx, ny;                                // global grid
    int row_base;                              // global row of array row 0
    int in_lo, in_hi;                          // global rows readable in the input set
    int out_lo, out_hi;                        // global rows this pass must produce
    int chunk_rows, nstrips, nchunks;
    // Edge chunks: the rows that need the careful kernel because of what lies ABOVE or BELOW them (PML rows, the grid's
    // first / last row, TFSF box rows, ghost rows and pushed rows of a slab) are a thin band, so the first and the last
    // chunk are cut just tall enough to hold it (first_rows / last_rows; 0 = uniform chunks) and everything between
    // is ordinary chunks of chunk_rows -- see chunk_span.
    int first_rows, last_rows;
    int cchunk_rows, ncchunks;                 // row partition of the SPECIAL strips (careful kernel): finer on small launches
    // col_fast: the special strips are special only for their COLUMNS (PML y-coefficients, the grid's first / last
    // column).  Where such a strip crosses an ordinary chunk every row is ordinary, so those items go to the column
    // variant of the warp-chain kernel (fd2d_chain.cu) and the careful kernel keeps the special strips only inside the
    // special chunks.
    int col_fast;
    // row_fast: likewise the special chunks that are special only for their ROWS (PML x-coefficients, the grid's first /
    // last row): where an ordinary strip crosses one, the item goes to the row variant of the warp-chain kernel and the
    // careful kernel keeps the corner blocks (special strip x special chunk) and the source cells.
    int row_fast;
    int tfsf, npml;
    const real *ezi_hist, *hxi_hist;           // [T][ny], [T][2]
    int src_i, src_j, src_hard;                // point source on dz (src_i < 0: none)
    int ident_row_lo, ident_row_hi, ident_col_lo, ident_col_hi;   // rows / cols [lo,hi) with identity PML coefficients
    int nf;                                    // fused running DFT: frequencies (0 = off, <= NFMAX)
    real *r_pt, *i_pt;                         // [nf][rows_alloc][ny] accumulators, updated in place by the owner warp
    long long dft_plane;                       // elements per frequency plane
    double dft_c[TMAX][NFMAX], dft_s[TMAX][NFMAX];   // phase factors of every sub-step
    // fused halo exchange over peer memory (multi-GPU): the rows within `halo` of the slab edges are ALSO stored into
    // the neighbours' ghost rows; completion is announced through flags in the neighbours' memory
    int push;                                  // this pass pushes its edge rows
    int halo_on, own_lo, own_hi;               // fused exchange enabled; rows this rank owns (host-side classification)
    int push_up_end, push_dn_begin;            // rows ro < push_up_end go up, rows ro >= push_dn_begin go down
    real *up_dz, *up_hx, *up_hy, *up_ihx, *up_ihy, *up_iz;   // neighbour above: its OUT-set arrays (NULL: none)
    real *dn_dz, *dn_hx, *dn_hy, *dn_ihx, *dn_ihy, *dn_iz;   // neighbour below
    long long up_shift, dn_shift;              // element offset of a global row in the neighbour's arrays minus mine
    int wait_flags, signal;                    // first / last pass of a call
    unsigned long long *sync_local;            // {flag written by up, flag written by down, counter[0], counter[1], error word}
    unsigned long long *flag_at_up, *flag_at_dn;   // where this rank announces itself (peer memory)
    unsigned long long epoch;                  // sequence number of this advance call (1, 2, ...)
    unsigned total_warps;                      // warps of the careful kernel of the pass (the only ones that touch ghosts)
    int write_ez;                              // 0: this pass leaves ez untouched (it is never read by a pass)
    // Lossy problems whose loss is local (a dielectric object in free space): outside rows [lz_row_lo, lz_row_hi) x
    // cols [lz_col_lo, lz_col_hi) nbz is 0 and iz is +0 in both state sets, where ez = naz*(dz-iz), iz += nbz*ez gives
    // the bits of ez = naz*dz and leaves iz alone.  Interior warps that stay outside the box then run the lossless
    // kernel (no iz / nbz traffic); the two interior kernels of a pass share one index space and each warp keeps or
    // drops itself by this box.
    int split_lossless;
    int lz_row_lo, lz_row_hi, lz_col_lo, lz_col_hi;
    int n_sstrips, n_schunks;                  // sorted ids of the strips / chunks the careful kernel owns
    int sstrips[MAX_SPECIAL], schunks[MAX_SPECIAL];
    int n_spairs;                              // single (strip, chunk) cells of otherwise ordinary strips and chunks that the
    int spairs[MAX_PAIRS][2];                  // careful kernel owns as well: the ones whose rows and columns see the point source
    double src[TMAX];
    unsigned long long negzero2;               // two float -0.0 (0x8000000080000000), opaque to the compiler: see pk_mul
    unsigned long long spin_ns;                // longest wait for a neighbour's flag before the pass gives up (halo_wait)
};

// ---- vector global access: V consecutive elements, naturally aligned
template <typename real, int V> struct VecIO;
template <> struct VecIO<float, 1> {
    static __device__ __forceinline__ void ld(const float *p, float (&d)[1]) { d[0] = __ldg(p); }
    static __device__ __forceinline__ void st(float *p, const float (&d)[1]) { *p = d[0]; }
};
template <> struct VecIO<float, 2> {
    static __device__ __forceinline__ void ld(const float *p, float (&d)[2]) {
        float2 v = __ldg(reinterpret_cast<const float2 *>(p)); d[0] = v.x; d[1] = v.y;
    }
    static __device__ __forceinline__ void st(float *p, const float (&d)[2]) {
        *reinterpret_cast<float2 *>(p) = make_float2(d[0], d[1]);
    }
};
template <> struct VecIO<float, 4> {
    static __device__ __forceinline__ void ld(const float *p, float (&d)[4]) {
        float4 v = __ldg(reinterpret_cast<const float4 *>(p)); d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
    static __device__ __forceinline__ void st(float *p, const float (&d)[4]) {
        *reinterpret_cast<float4 *>(p) = make_float4(d[0], d[1], d[2], d[3]);
    }
};
template <> struct VecIO<double, 1> {
    static __device__ __forceinline__ void ld(const double *p, double (&d)[1]) { d[0] = __ldg(p); }
    static __device__ __forceinline__ void st(double *p, const double (&d)[1]) { *p = d[0]; }
};
template <> struct VecIO<double, 2> {
    static __device__ __forceinline__ void ld(const double *p, double (&d)[2]) {
        double2 v = __ldg(reinterpret_cast<const double2 *>(p)); d[0] = v.x; d[1] = v.y;
    }
    static __device__ __forceinline__ void st(double *p, const double (&d)[2]) {
        *reinterpret_cast<double2 *>(p) = make_double2(d[0], d[1]);
    }
};

// ---- cp.async (LDGSTS): global -> shared without a register round trip; src_bytes = 0 zero-fills
template <int BYTES>
__device__ __forceinline__ void cp_async(void *smem_dst, const void *gsrc, int src_bytes) {
    const unsigned d = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    if (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
    else if (BYTES == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename real, int V>
__device__ __forceinline__ void lds_vec(const void *smem_src, real (&d)[V]) {
    if constexpr (sizeof(real) * V == 16) {
        const float4 t = *reinterpret_cast<const float4 *>(smem_src);
        const real *q = reinterpret_cast<const real *>(&t);
#pragma unroll
        for (int v = 0; v < V; ++v) d[v] = q[v];
    } else if constexpr (sizeof(real) * V == 8) {
        const float2 t = *reinterpret_cast<const float2 *>(smem_src);
        const real *q = reinterpret_cast<const real *>(&t);
#pragma unroll
        for (int v = 0; v < V; ++v) d[v] = q[v];
    } else {
        d[0] = *reinterpret_cast<const real *>(smem_src);
    }
}

// One grid row as it lives in registers.  A set is first the ARRIVING row of a stage (state at the stage's
// input time level; `ez` not yet meaningful), then the row the stage HOLDS (D/E advanced, H not yet), then --
// updated in place -- the row handed to the next stage.  Sets are never copied: with the row loop unrolled
// T+1 times the T+1 sets rotate through the roles under compile-time indices (no register moves).
template <typename real, int V>
struct RowSet {
    real dz[V], ez[V], hx[V], hy[V], ihx[V], ihy[V], naz[V], iz[V], nbz[V];
    real racc[NFMAX][V], iacc[NFMAX][V];       // running-DFT accumulators travelling with the row (DFT kernels only)
};

template <typename real, int V>
struct ColCoef {       // per-column PML coefficients and update masks, fixed for the whole march
    real gy2[V], gy3[V], fy1[V], fy2[V], fy3[V];
    unsigned dmask, hmask;         // bit v: D / H update applies to column jb+v
};

// One pipeline stage at sub-step s: finish D,E of the arriving row A (global row rs) and H of the held row Hd
// (global row rs-1), both in place.  FAST: interior warp -- no edge masks, no TFSF / source cells.
template <typename real, int V, int MODE, bool FAST, bool NAZR = false>
__device__ __forceinline__ void march_stage(const MarchParams<real> &p, const ColCoef<real, V> &c, RowSet<real, V> &A,
                                            RowSet<real, V> &Hd, const int rs, const int s, const int jb,
                                            const bool tf_cols, const bool src_cols, const void *naz_smem = nullptr,
                                            const void *naz_held_smem = nullptr) {
    constexpr bool LOSSY = (MODE & 1) != 0, DFT = (MODE & 2) != 0;
    static_assert(!NAZR || (FAST && !LOSSY), "the naz ring serves the plain interior kernel only");
    constexpr unsigned FULL = 0xffffffffu;
    const real half = real(0.5);
    const int hr = rs - 1;
    // FAST warps only touch rows / columns whose ten PML coefficients are the identity set (the host
    // guarantees it through fdtd2d_problem::ident_*): multiplications by exactly 1 are dropped -- an exact
    // identity for every input -- while 0*x is kept, because it decides the sign of a zero sum.
    const int rd = FAST ? rs : min(max(rs, 0), p.nx - 1);
    const int rh = FAST ? hr : min(max(hr, 0), p.nx - 1);
    const real gx2 = FAST ? real(1) : __ldg(p.gx2 + rd), gx3 = FAST ? real(1) : __ldg(p.gx3 + rd);
    const real fx1 = FAST ? real(0) : __ldg(p.fx1 + rh);
    const real fx2 = FAST ? real(1) : __ldg(p.fx2 + rh), fx3 = FAST ? real(1) : __ldg(p.fx3 + rh);
    const bool drow = FAST || ((rs >= 1) && (rs < p.nx));
    const bool hrow = FAST || ((hr >= 0) && (hr <= p.nx - 2));

    // ---- D of row rs:  dz = gx3*gy3*dz + gx2*gy2*0.5*(hy - hy[i-1] - hx + hx[j-1])
    const real hx_left = __shfl_up_sync(FULL, A.hx[V - 1], 1);
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const real hxl = (v == 0) ? hx_left : A.hx[v == 0 ? 0 : v - 1];
        const real curl = ((A.hy[v] - Hd.hy[v]) - A.hx[v]) + hxl;
        const real gy2 = FAST ? real(1) : c.gy2[v], gy3 = FAST ? real(1) : c.gy3[v];
        const real dn = ((gx3 * gy3) * A.dz[v]) + (((gx2 * gy2) * half) * curl);
        if (FAST) A.dz[v] = dn;
        else A.dz[v] = (drow && ((c.dmask >> v) & 1u)) ? dn : A.dz[v];
    }
    if (!FAST) {
        const int ia = p.npml - 1, iz_ = p.nx - p.npml, ja = p.npml - 1, jz = p.ny - p.npml;
        if (src_cols && rs == p.src_i) {             // point source (after the stencil, before inctdz)
#pragma unroll
            for (int v = 0; v < V; ++v)
                if (jb + v == p.src_j) A.dz[v] = fdtd::inject<real>(A.dz[v], p.src[s], p.src_hard);
        }
        if (tf_cols && rs >= ia && rs <= iz_) {      // inctdz: uses hxi of the previous step
            const real a = half * __ldg(p.hxi_hist + 2 * s), b = half * __ldg(p.hxi_hist + 2 * s + 1);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                if (jb + v == ja) A.dz[v] = A.dz[v] + a;
                if (jb + v == jz) A.dz[v] = A.dz[v] - b;
            }
        }
    }
    // ---- E of row rs
    real ezA[V], ezH[V];           // Ez of the arriving row (fresh) and of the held row (from the previous row trip)
    if constexpr (NAZR) {
        // deep pipelines: naz comes from its shared-memory ring and Ez is not kept in the row sets at all -- the
        // held row's Ez is the same product naz*dz evaluated again (same operands, same bits)
        real nzA[V], nzH[V];
        lds_vec<real, V>(naz_smem, nzA);
        lds_vec<real, V>(naz_held_smem, nzH);
#pragma unroll
        for (int v = 0; v < V; ++v) {
            ezA[v] = nzA[v] * A.dz[v];
            ezH[v] = nzH[v] * Hd.dz[v];
        }
    } else {
#pragma unroll
        for (int v = 0; v < V; ++v) {
            if (LOSSY) {
                A.ez[v] = A.naz[v] * (A.dz[v] - A.iz[v]);
                A.iz[v] = A.iz[v] + A.nbz[v] * A.ez[v];
            } else {
                A.ez[v] = A.naz[v] * A.dz[v];
            }
            ezA[v] = A.ez[v];
            ezH[v] = Hd.ez[v];
        }
    }
    if (DFT) {       // fourier of sub-step s on the fresh Ez: float64 product and sum, rounded into the array type
#pragma unroll
        for (int f = 0; f < NFMAX; ++f)
            if (f < p.nf) {
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const double e = static_cast<double>(A.ez[v]);
                    A.racc[f][v] = static_cast<real>(static_cast<double>(A.racc[f][v]) + p.dft_c[s][f] * e);
                    A.iacc[f][v] = static_cast<real>(static_cast<double>(A.iacc[f][v]) - p.dft_s[s][f] * e);
                }
            }
    }
    // ---- H of the held row hr (needs ez[hr][j+1] and ez[rs][j]), in place
    const real ez_right = __shfl_down_sync(FULL, ezH[0], 1);
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const real er = (v == V - 1) ? ez_right : ezH[v == V - 1 ? v : v + 1];
        const real cm = ezH[v] - er;
        const real cn = ezH[v] - ezA[v];
        const real sx = Hd.ihx[v] + cm;
        const real sy = Hd.ihy[v] + cn;
        const real fy1 = FAST ? real(0) : c.fy1[v], fy2 = FAST ? real(1) : c.fy2[v], fy3 = FAST ? real(1) : c.fy3[v];
        const real hx2 = (fy3 * Hd.hx[v]) + (fy2 * ((half * cm) + (fx1 * sx)));
        const real hy2 = (fx3 * Hd.hy[v]) - (fx2 * ((half * cn) + (fy1 * sy)));
        if (FAST) {
            Hd.ihx[v] = sx; Hd.ihy[v] = sy; Hd.hx[v] = hx2; Hd.hy[v] = hy2;
        } else {
            const bool up = hrow && ((c.hmask >> v) & 1u);
            Hd.ihx[v] = up ? sx : Hd.ihx[v];
            Hd.ihy[v] = up ? sy : Hd.ihy[v];
            Hd.hx[v] = up ? hx2 : Hd.hx[v];
            Hd.hy[v] = up ? hy2 : Hd.hy[v];
        }
    }
    if (!FAST && p.tfsf) {
        const int ia = p.npml - 1, iz_ = p.nx - p.npml, ja = p.npml - 1, jz = p.ny - p.npml;
        if (tf_cols && hr >= ia && hr <= iz_) {      // incthx
            const real *ez_i = p.ezi_hist + (size_t)s * p.ny;
            const real a = half * __ldg(ez_i + ja), b = half * __ldg(ez_i + jz);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                if (jb + v == ja - 1) Hd.hx[v] = Hd.hx[v] + a;
                if (jb + v == jz) Hd.hx[v] = Hd.hx[v] - b;
            }
        }
        if (hr == ia - 1 || hr == iz_) {             // incthy (two rows of the whole grid)
            const real *ez_i = p.ezi_hist + (size_t)s * p.ny;
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const int j = jb + v;
                if (j >= ja && j <= jz) {
                    const real h = half * __ldg(ez_i + j);
                    if (hr == ia - 1) Hd.hy[v] = Hd.hy[v] - h;
                    if (hr == iz_) Hd.hy[v] = Hd.hy[v] + h;
                }
            }
        }
    }
}

// ---- packed fp32 (sm_100 FADD2 / FFMA2: two IEEE operations per instruction, half the issue slots and code size).
// Each half rounds exactly like the scalar instruction, so results stay bit-identical to the reference order.  One trap:
// ptxas contracts a packed multiply with a following packed add into FFMA2 even with --fmad=false (observed with
// CUDA 12.9: __fmul2_rn + __fadd2_rn -> one FFMA2), which would drop a rounding.  A product is therefore issued as
// FFMA2(a, b, -0.0) with the -0.0 pair taken from a kernel parameter the compiler cannot see through:
// a*b + (-0) rounds once, to exactly RN(a*b) (signed zeros included), and an FFMA2 cannot absorb the next add.
__device__ __forceinline__ float2 pk_add(const float2 a, const float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 pk_sub(const float2 a, const float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 pk_mul(const float2 a, const float2 b, const float2 negzero) { return __ffma2_rn(a, b, negzero); }

// Interior stage in packed arithmetic (float, even V): same operations in the same order as march_stage<.., FAST>.
template <int V, bool NAZR, bool LOSSY>
__device__ __forceinline__ void march_stage_pk(RowSet<float, V> &A, RowSet<float, V> &Hd, const float2 negzero,
                                               const void *naz_smem, const void *naz_held_smem) {
    static_assert(V % 2 == 0, "packed stage needs column pairs");
    static_assert(!(NAZR && LOSSY), "the naz ring serves the plain interior kernel only");
    constexpr unsigned FULL = 0xffffffffu;
    const float2 half2 = make_float2(0.5f, 0.5f), zero2 = make_float2(0.f, 0.f);
    // ---- D of the arriving row: dz = dz + 0.5*(((hy - hy[i-1]) - hx) + hx[j-1])
    const float hx_left = __shfl_up_sync(FULL, A.hx[V - 1], 1);
#pragma unroll
    for (int v = 0; v < V; v += 2) {
        const float2 a1 = pk_sub(make_float2(A.hy[v], A.hy[v + 1]), make_float2(Hd.hy[v], Hd.hy[v + 1]));
        const float2 a2 = pk_sub(a1, make_float2(A.hx[v], A.hx[v + 1]));
        const float2 curl = make_float2(a2.x + (v == 0 ? hx_left : A.hx[v == 0 ? 0 : v - 1]), a2.y + A.hx[v]);   // shifted pair: scalar
        const float2 dn = pk_add(make_float2(A.dz[v], A.dz[v + 1]), pk_mul(half2, curl, negzero));
        A.dz[v] = dn.x; A.dz[v + 1] = dn.y;
    }
    // ---- E of both rows
    float ezA[V], ezH[V];
    if constexpr (NAZR) {
        float nzA[V], nzH[V];
        lds_vec<float, V>(naz_smem, nzA);
        lds_vec<float, V>(naz_held_smem, nzH);
#pragma unroll
        for (int v = 0; v < V; v += 2) {
            const float2 a = pk_mul(make_float2(nzA[v], nzA[v + 1]), make_float2(A.dz[v], A.dz[v + 1]), negzero);
            const float2 h = pk_mul(make_float2(nzH[v], nzH[v + 1]), make_float2(Hd.dz[v], Hd.dz[v + 1]), negzero);
            ezA[v] = a.x; ezA[v + 1] = a.y; ezH[v] = h.x; ezH[v + 1] = h.y;
        }
    } else {
#pragma unroll
        for (int v = 0; v < V; v += 2) {
            float2 a;
            if constexpr (LOSSY) {      // ez = naz*(dz - iz); iz = iz + nbz*ez
                const float2 iz = make_float2(A.iz[v], A.iz[v + 1]);
                a = pk_mul(make_float2(A.naz[v], A.naz[v + 1]), pk_sub(make_float2(A.dz[v], A.dz[v + 1]), iz), negzero);
                const float2 i2 = pk_add(iz, pk_mul(make_float2(A.nbz[v], A.nbz[v + 1]), a, negzero));
                A.iz[v] = i2.x; A.iz[v + 1] = i2.y;
            } else {
                a = pk_mul(make_float2(A.naz[v], A.naz[v + 1]), make_float2(A.dz[v], A.dz[v + 1]), negzero);
            }
            A.ez[v] = a.x; A.ez[v + 1] = a.y;
            ezA[v] = a.x; ezA[v + 1] = a.y; ezH[v] = Hd.ez[v]; ezH[v + 1] = Hd.ez[v + 1];
        }
    }
    // ---- H of the held row: ihx += cm; ihy += cn; hx = hx + (0.5*cm + 0*ihx); hy = hy - (0.5*cn + 0*ihy)
    const float ez_right = __shfl_down_sync(FULL, ezH[0], 1);
#pragma unroll
    for (int v = 0; v < V; v += 2) {
        const float2 e = make_float2(ezH[v], ezH[v + 1]);
        const float2 cm = make_float2(ezH[v] - ezH[v + 1], ezH[v + 1] - (v + 2 < V ? ezH[v + 2 < V ? v + 2 : v] : ez_right));   // shifted pair: scalar
        const float2 cn = pk_sub(e, make_float2(ezA[v], ezA[v + 1]));
        const float2 sx = pk_add(make_float2(Hd.ihx[v], Hd.ihx[v + 1]), cm);
        const float2 sy = pk_add(make_float2(Hd.ihy[v], Hd.ihy[v + 1]), cn);
        const float2 tx = pk_add(pk_mul(half2, cm, negzero), pk_mul(zero2, sx, negzero));
        const float2 ty = pk_add(pk_mul(half2, cn, negzero), pk_mul(zero2, sy, negzero));
        const float2 hx2 = pk_add(make_float2(Hd.hx[v], Hd.hx[v + 1]), tx);
        const float2 hy2 = pk_sub(make_float2(Hd.hy[v], Hd.hy[v + 1]), ty);
        Hd.ihx[v] = sx.x; Hd.ihx[v + 1] = sx.y; Hd.ihy[v] = sy.x; Hd.ihy[v + 1] = sy.y;
        Hd.hx[v] = hx2.x; Hd.hx[v + 1] = hx2.y; Hd.hy[v] = hy2.x; Hd.hy[v + 1] = hy2.y;
    }
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// k-th id (0-based) of the ascending sequence 0,1,2,... with the sorted ids in `skip` removed
__device__ __forceinline__ int kth_not_in(int k, const int *skip, int n) {
    for (int q = 0; q < n; ++q)
        if (skip[q] <= k) ++k;
    return k;
}


// rows [i0, i1) of chunk k of the pass (uniform chunks, or a short first and last chunk around uniform ones)
template <typename real>
__host__ __device__ __forceinline__ void chunk_span(const MarchParams<real> &p, const int k, int &i0, int &i1) {
    if (p.first_rows == 0) {
        i0 = p.out_lo + k * p.chunk_rows;
        i1 = i0 + p.chunk_rows < p.out_hi ? i0 + p.chunk_rows : p.out_hi;
    } else if (k == 0) {
        i0 = p.out_lo;
        i1 = p.out_lo + p.first_rows;
    } else if (k == p.nchunks - 1) {
        i0 = p.out_hi - p.last_rows;
        i1 = p.out_hi;
    } else {
        const int end = p.out_hi - p.last_rows;
        i0 = p.out_lo + p.first_rows + (k - 1) * p.chunk_rows;
        i1 = i0 + p.chunk_rows < end ? i0 + p.chunk_rows : end;
    }
}

// ---- which (strip, rows) a warp owns.  Two kinds of launch share one index space per pass:
//   FAST    : interior warps -- (ordinary strips) x (ordinary chunks): every column (halo included) is an ordinary
//             cell and every row touched (warm-up, drain and unroll overrun included) is an ordinary stored row;
//   careful : the listed special strips x all rows (in their own, finer row partition), plus ordinary strips x the
//             listed special chunks, plus the single (strip, chunk) cells that see the point source; or everything
//             (all_careful).
// `lossy_kernel`: with the lossless-outside split two interior kernels run over the FAST index space and each warp
// keeps or drops itself by the promised box.  Returns false when the warp has nothing to do.
template <bool FAST, typename real>
__device__ __forceinline__ bool decode_item(const MarchParams<real> &p, const int w, const int all_careful, const int V,
                                            const int T, const bool lossy_kernel, int &strip, int &i0, int &i1) {
    const int nsf = p.nstrips - p.n_sstrips, ncf = p.nchunks - p.n_schunks;   // ordinary strips / chunks
    int chunk;
    bool own_partition = false;                 // special strips: their own (finer, uniform) row partition
    if (FAST) {
        if (w >= nsf * ncf) return false;
        strip = kth_not_in(w % nsf, p.sstrips, p.n_sstrips);
        chunk = kth_not_in(w / nsf, p.schunks, p.n_schunks);
        for (int q = 0; q < p.n_spairs; ++q)
            if (strip == p.spairs[q][0] && chunk == p.spairs[q][1]) return false;     // the careful kernel has this one
        if (p.split_lossless) {
            // rows and columns this warp touches (warm-up, drain and fetch run-ahead included), as the host classifies them
            const int W = 32 * V, HALO = ((T + V - 1) / V) * V, USE = W - 2 * HALO;
            const int c0 = strip * USE - HALO, c1 = c0 + W;
            int r0, r1;
            chunk_span(p, chunk, r0, r1);
            const int lo = r0 - T - 1, hi = r1 + 2 * T + RING + 2;
            const bool in_box = c0 < p.lz_col_hi && c1 > p.lz_col_lo && lo < p.lz_row_hi && hi > p.lz_row_lo;
            if (in_box != lossy_kernel) return false;    // lossy kernel: warps meeting the box; lossless kernel: the others
        }
    } else if (all_careful) {
        if (w >= p.nstrips * p.nchunks) return false;
        strip = w % p.nstrips;
        chunk = w / p.nstrips;
    } else {
        // special strips: all rows, in their own (finer) row partition -- or, with col_fast, the special chunks only
        const int na = p.n_sstrips * (p.col_fast ? p.n_schunks : p.ncchunks);
        if (w < na) {
            strip = p.sstrips[w % p.n_sstrips];
            if (p.col_fast) {
                chunk = p.schunks[w / p.n_sstrips];
            } else {
                chunk = w / p.n_sstrips;
                own_partition = true;
            }
        } else {
            const int x = w - na, nb = p.row_fast ? 0 : nsf * p.n_schunks;
            if (x < nb) {
                strip = kth_not_in(x % nsf, p.sstrips, p.n_sstrips);
                chunk = p.schunks[x / nsf];
            } else {
                if (x - nb >= p.n_spairs) return false;
                strip = p.spairs[x - nb][0];
                chunk = p.spairs[x - nb][1];
            }
        }
    }
    if (own_partition) {
        i0 = p.out_lo + chunk * p.cchunk_rows;
        i1 = min(i0 + p.cchunk_rows, p.out_hi);
    } else {
        chunk_span(p, chunk, i0, i1);
    }
    return true;
}

// ---- fused halo exchange (multi-GPU).  Ghost rows are read, and edge rows pushed, by careful warps only (the host
// lists those chunks as special), so the handshake lives in the careful kernels; the interior kernels carry none.
// The wait is bounded in TIME: a neighbour that never arrives (a rank that died or skipped a call) makes the warp give
// up after spin_ns nanoseconds (default 20 s), raise the error word sync_local[4] = epoch*4 + side and go on -- the
// pass then ends (with stale ghost rows) and the host reports the error (fdtd2d_halo_status) instead of hanging forever.
constexpr unsigned long long HALO_SPIN_NS = 20ull * 1000 * 1000 * 1000;

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <typename real>
__device__ __forceinline__ void halo_wait(const MarchParams<real> &p, const int lane) {
    if (!p.wait_flags) return;
    // neighbours must have finished their previous call: their pushes into my ghost rows have landed, and they no
    // longer read the ghost rows this call's pushes will overwrite
    if (lane == 0) {
        const unsigned long long need = p.epoch - 1;
        for (int side = 0; side < 2; ++side) {
            if ((side == 0 ? p.flag_at_up : p.flag_at_dn) == nullptr) continue;
            unsigned polls = 0;
            unsigned long long t0 = 0;
            while (ld_acquire_sys(p.sync_local + side) < need) {
                if ((++polls & 1023u) != 0) continue;
                const unsigned long long now = global_timer_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > p.spin_ns) {
                    atomicExch(p.sync_local + 4, (p.epoch << 2) | (unsigned long long)(side + 1));
                    break;
                }
            }
        }
    }
    __syncwarp();
}

template <typename real>
__device__ __forceinline__ void halo_signal(const MarchParams<real> &p, const int lane) {
    if (!p.signal) return;
    // the last careful warp of the pass announces completion to the neighbours
    __threadfence_system();
    __syncwarp();
    if (lane == 0) {
        unsigned long long *counter = p.sync_local + 2 + (p.epoch & 1ull);
        const unsigned long long done = atomicAdd(counter, 1ull);
        if (done + 1 == (unsigned long long)p.total_warps) {
            atomicExch(counter, 0ull);
            __threadfence_system();
            if (p.flag_at_up != nullptr) st_release_sys(p.flag_at_up, p.epoch);
            if (p.flag_at_dn != nullptr) st_release_sys(p.flag_at_dn, p.epoch);
        }
    }
}

// peer stores of one finished row (global row ro, element offset off_s in MY arrays) into the neighbours' ghost rows
template <typename real, int V, bool LOSSY>
__device__ __forceinline__ void push_row(const MarchParams<real> &p, const long long off_s, const int ro,
                                         const real (&dz)[V], const real (&hx)[V], const real (&hy)[V],
                                         const real (&ihx)[V], const real (&ihy)[V], const real (&iz)[V]) {
    if (ro < p.push_up_end && p.up_dz != nullptr) {
        const long long o = off_s + p.up_shift;
        VecIO<real, V>::st(p.up_dz + o, dz);   VecIO<real, V>::st(p.up_hx + o, hx);
        VecIO<real, V>::st(p.up_hy + o, hy);   VecIO<real, V>::st(p.up_ihx + o, ihx);
        VecIO<real, V>::st(p.up_ihy + o, ihy);
        if (LOSSY) VecIO<real, V>::st(p.up_iz + o, iz);
    }
    if (ro >= p.push_dn_begin && p.dn_dz != nullptr) {
        const long long o = off_s + p.dn_shift;
        VecIO<real, V>::st(p.dn_dz + o, dz);   VecIO<real, V>::st(p.dn_hx + o, hx);
        VecIO<real, V>::st(p.dn_hy + o, hy);   VecIO<real, V>::st(p.dn_ihx + o, ihx);
        VecIO<real, V>::st(p.dn_ihy + o, ihy);
        if (LOSSY) VecIO<real, V>::st(p.dn_iz + o, iz);
    }
}

// ---- host side, shared by the two translation units (defined in fd2d_march.cu)
struct Tuning {
    int force_v = 0, chunk_rows = 0, warps = 0, careful = 0;
    int split = 1;               // 0 = ignore the lossless-outside promise (tests: the lossy kernel everywhere)
    int serial = 2;              // 2 = fork the edge kernel onto a side stream (measured +2 %); 1 = edge then interior in order;
                                 // 3 (deep passes) = interior first, the edge kernel backfills at one warp per CTA
    int variant = 0;             // kernel-shape experiments of the deep passes (0 = the shipped shape)
    int deep = 1;                // 0 = never use the deep passes of fd2d_deep.cu; 2 = the smem-resident careful kernel at every depth
    int edge_chunks = 1;         // 1 = short first / last chunk around the rows that need the careful kernel; 0 = uniform chunks
    int col_fast = 1;            // 1 = PML-column strips x ordinary chunks through the warp-chain kernel's column variant
    int row_fast = 1;            // 1 = ordinary strips x PML-row chunks through its row variant
    unsigned long long spin_ns = HALO_SPIN_NS;
};
extern Tuning g_tune;

struct SideStream { cudaStream_t stream; cudaEvent_t fork, join; cudaStream_t backfill; };   // backfill: default priority
// one high-priority side stream + fork/join events per (device, launch stream), created on first use
SideStream *side_stream(cudaStream_t launch);

struct PassCounts { bool all_careful; int n_fast, n_careful, n_col, n_row; };   // n_col / n_row: the items of col_fast / row_fast

// Classify strips and chunks on the host (same conditions as the kernels rely on): fills the special lists, the single
// source cells, the careful row partition and total_warps of `mp` for a pass of vector width V and depth T.
template <typename real>
PassCounts classify_pass(MarchParams<real> &mp, const int V, const int T) {
    const int W = 32 * V, HALO = ((T + V - 1) / V) * V, USE = W - 2 * HALO;
    const int ja = mp.npml - 1, jz = mp.ny - mp.npml, ia = mp.npml - 1, iz_ = mp.nx - mp.npml;
    int ns = 0, nc = 0;
    bool overflow = g_tune.careful != 0;
    for (int k = 0; k < mp.nstrips && !overflow; ++k) {
        const int c0 = k * USE - HALO, c1 = c0 + W;     // columns [c0, c1)
        bool special = (c0 < max(1, mp.ident_col_lo)) || (c1 > min(mp.ny - 1, mp.ident_col_hi));
        if (mp.tfsf) special = special || (ja - 1 >= c0 && ja - 1 < c1) || (ja >= c0 && ja < c1) || (jz >= c0 && jz < c1);
        if (special) {
            if (ns == MAX_SPECIAL) overflow = true;
            else mp.sstrips[ns++] = k;
        }
    }
    for (int k = 0; k < mp.nchunks && !overflow; ++k) {
        int i0, i1;
        chunk_span(mp, k, i0, i1);
        const int lo = i0 - T - 1, hi = i1 + 2 * T + RING + 2; // rows touched (fetch run-ahead included): [lo, hi)
        bool special = (lo < max(max(1, mp.in_lo), mp.ident_row_lo)) || (hi > min(min(mp.nx - 1, mp.in_hi), mp.ident_row_hi));
        if (mp.tfsf) special = special || (ia - 1 >= lo && ia - 1 < hi) || (iz_ >= lo && iz_ < hi);
        // fused halo exchange: chunks that touch a ghost row or own a pushed row carry the handshake
        if (mp.halo_on) special = special || lo < mp.own_lo || hi > mp.own_hi || i0 < mp.push_up_end || i1 > mp.push_dn_begin;
        if (special) {
            if (nc == MAX_SPECIAL) overflow = true;
            else mp.schunks[nc++] = k;
        }
    }
    // The point source is ONE cell: only the (strip, chunk) cells whose columns and rows see it go to the careful kernel,
    // not its whole strip and its whole chunk.
    int np = 0;
    if (mp.src_i >= 0 && !overflow) {
        auto listed = [](const int *a, int n, int k) { for (int q = 0; q < n; ++q) if (a[q] == k) return true; return false; };
        for (int k = 0; k < mp.nstrips && !overflow; ++k) {
            const int c0 = k * USE - HALO, c1 = c0 + W;
            if (!(mp.src_j >= c0 && mp.src_j < c1) || listed(mp.sstrips, ns, k)) continue;
            for (int c = 0; c < mp.nchunks && !overflow; ++c) {
                int i0, i1;
                chunk_span(mp, c, i0, i1);
                const int lo = i0 - T - 1, hi = i1 + 2 * T + RING + 2;
                if (!(mp.src_i >= lo && mp.src_i < hi) || listed(mp.schunks, nc, c)) continue;
                if (np == MAX_PAIRS) overflow = true;
                else { mp.spairs[np][0] = k; mp.spairs[np][1] = c; ++np; }
            }
        }
    }
    mp.n_spairs = overflow ? 0 : np;
    PassCounts out;
    if (overflow) {                                      // tiny grids (or a forced careful run): everything through the careful kernel
        mp.n_sstrips = mp.n_schunks = 0;
        mp.cchunk_rows = mp.chunk_rows; mp.ncchunks = mp.nchunks;
        mp.total_warps = (unsigned)(mp.nstrips * mp.nchunks);
        out.all_careful = true; out.n_fast = 0; out.n_careful = mp.nstrips * mp.nchunks; out.n_col = out.n_row = 0;
        mp.col_fast = mp.row_fast = 0;
        return out;
    }
    mp.n_sstrips = ns; mp.n_schunks = nc;
    const int nsf = mp.nstrips - ns, ncf = mp.nchunks - nc;
    const int n_fast = nsf * ncf;                        // (np of them exit at once)
    // A careful warp is several times slower per row than an interior warp, and a launch cannot end before its slowest
    // warp.  On big launches (tens of interior waves) that is hidden; on small ones (row blocks of a streamed run, small
    // grids) the special strips get a finer row partition so that their warps finish with the interior's.
    const int rows = mp.out_hi - mp.out_lo;
    const bool small_launch = n_fast < 16 * MAX_WARPS * fdtd::sm_count();
    mp.cchunk_rows = small_launch ? max(1, min(mp.chunk_rows, max(4 * T, 32))) : mp.chunk_rows;
    mp.ncchunks = (rows + mp.cchunk_rows - 1) / mp.cchunk_rows;
    // col_fast needs a source cell outside the special strips (their items carry no source code)
    if (mp.col_fast && mp.src_i >= 0)
        for (int q = 0; q < ns; ++q) {
            const int c0 = mp.sstrips[q] * USE - HALO;
            if (mp.src_j >= c0 && mp.src_j < c0 + W) mp.col_fast = 0;
        }
    if (ns == 0 || ncf == 0) mp.col_fast = 0;
    // row_fast needs a source cell outside the rows the special chunks touch (their items carry no source code); without
    // col_fast the special strips run through ALL rows in their own partition, corner blocks included, so nothing else changes
    if (mp.row_fast && mp.src_i >= 0)
        for (int q = 0; q < nc; ++q) {
            int i0, i1;
            chunk_span(mp, mp.schunks[q], i0, i1);
            if (mp.src_i >= i0 - T - 1 && mp.src_i < i1 + 2 * T + RING + 2) mp.row_fast = 0;
        }
    if (nc == 0 || nsf == 0) mp.row_fast = 0;
    out.all_careful = false; out.n_fast = n_fast;
    out.n_col = mp.col_fast ? ns * ncf : 0;
    out.n_row = mp.row_fast ? nsf * nc : 0;
    out.n_careful = (mp.col_fast ? ns * nc : ns * mp.ncchunks) + (mp.row_fast ? 0 : nsf * nc) + np;
    mp.total_warps = (unsigned)out.n_careful;
    return out;
}

// deep passes (fd2d_deep.cu): float, 4-wide vectors, depth 8 or 12, no fused DFT
bool deep_supported(int T, bool lossy);
int launch_march_deep(MarchParams<float> &mp, int T, bool lossy, cudaStream_t st);
// the shared-memory-resident careful kernel (any depth <= TMAX), used by the deep passes
int launch_careful2(MarchParams<float> &mp, int T, bool lossy, int items, int all_careful, cudaStream_t st, int V = 4, int warps_per_cta = 0);
void preload_deep(bool lossy);
// warp-chain passes (fd2d_chain.cu): the interior items of a depth-8 / depth-12 pass as a TMA-fed pipeline of warps
bool chain_supported(int T, bool lossy);
int launch_march_chain(const MarchParams<float> &mp, int T, int shape, int items, cudaStream_t st, int kind = 0);   // kind: 0 interior, 1 column items, 2 row items
void preload_chain();

}  // namespace fdtd_march
