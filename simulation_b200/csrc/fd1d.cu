// 1D Ex/Hy line: the reference-named step functions (one kernel each) and a fused, temporally blocked
// advance.  Update rules, ranges and evaluation order follow the reference numpy programs
// (fd1d/program/fd1d_1_5.py:63-72 FDTD form, fd1d_2_3.py:73-94 flux/Debye form); no FMA contraction.
//
// Fused path: every WARP owns a segment, keeps it plus T halo cells per side in registers (consecutive cells per
// lane, neighbours by warp shuffle), runs T leap-frog steps on chip and writes its segment to the other
// (ping-pong) array set -- 2*T-cell recompute per warp instead of a grid-wide barrier per step.
#include "common.cuh"

namespace {

using fdtd::inject;

// --------------------------------------------------------------------------- reference-named kernels
template <typename real>
__global__ void k1_exfield(int nx, const real *ca, const real *cb, real *ex, const real *hy, int src_index,
                           int src_hard, double src_value) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nx) return;
    real e = ex[i];
    if (i >= 1) {
        const real a = ca ? ca[i] : real(1), b = cb ? cb[i] : real(0.5);
        e = (a * e) + (b * (hy[i - 1] - hy[i]));
    }
    if (i == src_index) e = inject<real>(e, src_value, src_hard);
    ex[i] = e;
}

template <typename real>
__global__ void k1_dxfield(int nx, real *dx, const real *hy, int src_index, int src_hard, double src_value) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nx) return;
    real d = dx[i];
    if (i >= 1) d = d + real(0.5) * (hy[i - 1] - hy[i]);
    if (i == src_index) d = inject<real>(d, src_value, src_hard);
    dx[i] = d;
}

template <typename real>
__global__ void k1_exfield_flux(int nx, const real *nax, const real *nbx, const real *ncx, const real *ndx,
                                const real *dx, real *ix, real *sx, real *ex) {
    int i = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nx) return;
    real e;
    if (sx) {
        const real c = ncx[i], s = sx[i];
        e = nax[i] * ((dx[i] - ix[i]) - (c * s));
        sx[i] = (c * s) + (ndx[i] * e);
    } else {
        e = nax[i] * (dx[i] - ix[i]);
    }
    ex[i] = e;
    ix[i] = ix[i] + nbx[i] * e;
}

// the ABC must see the finished E half step and finish before the H half step -> separate tiny kernel
template <typename real>
__global__ void k1_abc(int nx, real *ex, real *bc) {
    real e1 = ex[1], b0 = bc[0], b1 = bc[1];
    ex[0] = b0; bc[0] = b1; bc[1] = e1;
    real e2 = ex[nx - 2], b3 = bc[3], b2 = bc[2];
    ex[nx - 1] = b3; bc[3] = b2; bc[2] = e2;
}

template <typename real>
__global__ void k1_hy(int nx, const real *ex, real *hy) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nx - 1) hy[i] = hy[i] + real(0.5) * (ex[i] - ex[i + 1]);
}

// --------------------------------------------------------------------------------- fused advance
// One WARP owns a segment of the line and marches it T steps forward entirely in registers: every lane keeps KC
// CONSECUTIVE cells (Ex, Hy, the pointwise state and the coefficients), so a cell's neighbours are registers of
// the same lane except at the lane's two ends, which take one warp shuffle per half step (hy[i-1] from the lane
// below, ex[i+1] from the lane above).  No shared memory, no barrier; warps are independent and recompute T halo
// cells per side (the 1D twin of the 2D march kernel).  A pass = one launch = T steps = one trip of the state
// through memory; the ping-pong array sets make the warps order-free.
constexpr int T1MAX = 64;     // most steps per call of the time-block argument (deeper requests are split)

// DFT: the running-DFT variant also carries 2 x nf accumulators per cell in registers -> half the cells per lane
template <typename real, bool DFT = false> struct LineShape {
    static constexpr int KC = (sizeof(real) == 4 ? 16 : 8) / (DFT ? 2 : 1);   // cells per lane (registers: 9 arrays x KC x words)
    static constexpr int VEC = 16 / (int)sizeof(real);        // cells per 16-byte vector access
    static constexpr int W = 32 * KC;                         // cells staged per warp
    static constexpr int TMAX = 2 * KC;                       // deepest pass: halo <= W/8 per side
};

template <typename real>
struct LineParams {
    const real *in[5];        // ex, hy, dx, ix, sx
    real *out[5];
    const real *bc_in;
    real *bc_out;
    const real *ca, *cb, *nax, *nbx, *ncx, *ndx;
    int nx, T, halo, useful;  // halo = T rounded up to whole vectors; useful = W - 2*halo cells produced per warp
    int nwarps;
    int abc, src_field, src_index, src_hard;
    double src[T1MAX];        // waveform samples of this pass (by value: no staging buffer to manage)
    unsigned long long negzero2;   // two float -0.0, opaque to the compiler (packed products, see fd2d_march.cu pk_mul)
};

// Running DFT carried through a pass (programs 2_2 / 2_3, fd1d/program/fd1d_2_2.py:65-71): every step samples Ex
// between the E update and the ABC.  Phase factors of the pass by value (float64, evaluated on the host with the
// reference's expression); accumulators are read and written in place by the lane that owns the cell.
constexpr int DFT_NF = 3;         // frequencies the fused variant carries (more: unfused per-step path on the host side)
constexpr int DFT_TMAX = 16;      // deepest DFT pass (= LineShape<float, true>::TMAX)
template <typename real>
struct LineDft {
    real *r_pt, *i_pt, *r_in, *i_in;      // r_pt / i_pt: nf x nx; r_in / i_in: nf (may be NULL)
    int nf, sample;                       // sample: cell whose Ex feeds r_in / i_in (10 in the reference)
    double c[DFT_NF][DFT_TMAX], s[DFT_NF][DFT_TMAX];
};

// packed fp32 (FADD2 / FFMA2): each half rounds like the scalar instruction; a product is FFMA2(a, b, -0.0) with the
// -0.0 pair from a kernel parameter, because ptxas contracts a packed multiply + add even under --fmad=false
__device__ __forceinline__ float2 pk_add(const float2 a, const float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 pk_mul(const float2 a, const float2 b, const float2 nz) { return __ffma2_rn(a, b, nz); }

template <typename real, int KC>
__device__ __forceinline__ void ld_cells(const real *base, long long g0, bool vec_ok, int nx, real (&d)[KC], real fill) {
    constexpr int VEC = 16 / (int)sizeof(real);
    if (vec_ok) {
#pragma unroll
        for (int v = 0; v < KC / VEC; ++v) {
            const float4 t = __ldg(reinterpret_cast<const float4 *>(base + g0 + v * VEC));
            const real *q = reinterpret_cast<const real *>(&t);
#pragma unroll
            for (int e = 0; e < VEC; ++e) d[v * VEC + e] = q[e];
        }
    } else {
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            const long long g = g0 + k;
            d[k] = (g >= 0 && g < nx) ? __ldg(base + g) : fill;
        }
    }
}

// EDGE: the staged range touches an end of the line (cells outside the line, the never-updated ex[0], ex[nx-1],
// hy[nx-1], the ABC) or is not vector-aligned; interior warps run without any mask.
template <typename real, bool FLUX, bool DEBYE, bool EDGE, bool DFT = false>
__device__ __forceinline__ void line_body(const LineParams<real> &p, const int w, const int lane,
                                          const LineDft<real> *dft = nullptr) {
    using Shape = LineShape<real, DFT>;
    constexpr int KC = Shape::KC, VEC = Shape::VEC, W = Shape::W;
    constexpr unsigned FULL = 0xffffffffu;
    const real half = real(0.5);
    const int seg_lo = w * p.useful, seg_hi = min(seg_lo + p.useful, p.nx);
    const int base = seg_lo - p.halo;              // global cell of the warp's first staged cell (may be negative)
    const int g0 = base + lane * KC;               // first cell of this lane
    const bool vec_ok = !EDGE;                     // interior warps: every staged cell exists and rows are 16-byte aligned

    real ex[KC], hy[KC], dx[KC], ix[KC], sx[KC], c0[KC], c1[KC], c2[KC], c3[KC];
#pragma unroll
    for (int k = 0; k < KC; ++k) dx[k] = ix[k] = sx[k] = c2[k] = c3[k] = real(0);
    ld_cells<real, KC>(p.in[0], g0, vec_ok, p.nx, ex, real(0));
    ld_cells<real, KC>(p.in[1], g0, vec_ok, p.nx, hy, real(0));
    if (FLUX) {
        ld_cells<real, KC>(p.in[2], g0, vec_ok, p.nx, dx, real(0));
        ld_cells<real, KC>(p.in[3], g0, vec_ok, p.nx, ix, real(0));
        ld_cells<real, KC>(p.nax, g0, vec_ok, p.nx, c0, real(1));
        ld_cells<real, KC>(p.nbx, g0, vec_ok, p.nx, c1, real(0));
        if (DEBYE) {
            ld_cells<real, KC>(p.in[4], g0, vec_ok, p.nx, sx, real(0));
            ld_cells<real, KC>(p.ncx, g0, vec_ok, p.nx, c2, real(0));
            ld_cells<real, KC>(p.ndx, g0, vec_ok, p.nx, c3, real(0));
        }
    } else {
        if (p.ca) ld_cells<real, KC>(p.ca, g0, vec_ok, p.nx, c0, real(1));
        else {
#pragma unroll
            for (int k = 0; k < KC; ++k) c0[k] = real(1);
        }
        if (p.cb) ld_cells<real, KC>(p.cb, g0, vec_ok, p.nx, c1, real(0.5));
        else {
#pragma unroll
            for (int k = 0; k < KC; ++k) c1[k] = real(0.5);
        }
    }
    const int src_k = (p.src_index >= g0 && p.src_index < g0 + KC) ? p.src_index - g0 : -1;   // -1: not in this lane

    // ABC state (EDGE warps whose staged range contains an end of the line; every such warp runs its own copy)
    const bool has_left = EDGE && p.abc && (base <= 0);
    const bool has_right = EDGE && p.abc && (base + W >= p.nx);
    const int k_first = (has_left && 0 >= g0 && 0 < g0 + KC) ? -g0 : -1;                       // my index of cell 0
    const int k_last = (has_right && p.nx - 1 >= g0 && p.nx - 1 < g0 + KC) ? p.nx - 1 - g0 : -1;   // ... of cell nx-1
    real b0 = real(0), b1 = real(0), b2 = real(0), b3 = real(0);
    if (EDGE && p.abc) { b0 = p.bc_in[0]; b1 = p.bc_in[1]; b2 = p.bc_in[2]; b3 = p.bc_in[3]; }

    // running-DFT accumulators of the cells this lane OWNS (plain loads: other warps write their own cells of the same
    // arrays during this launch); the lane owning the sample cell also carries r_in / i_in
    constexpr int NFR = DFT ? DFT_NF : 1, KCR = DFT ? KC : 1;
    real acc_r[NFR][KCR], acc_i[NFR][KCR], in_r[NFR], in_i[NFR];
    int smp_k = -1;
    if constexpr (DFT) {
#pragma unroll
        for (int f = 0; f < DFT_NF; ++f) {
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                const int g = g0 + k;
                const bool mine = f < dft->nf && g >= seg_lo && g < seg_hi;
                acc_r[f][k] = mine ? dft->r_pt[(size_t)f * p.nx + g] : real(0);
                acc_i[f][k] = mine ? dft->i_pt[(size_t)f * p.nx + g] : real(0);
            }
        }
        if (dft->r_in && dft->sample >= max(g0, seg_lo) && dft->sample < min(g0 + KC, seg_hi)) smp_k = dft->sample - g0;
#pragma unroll
        for (int f = 0; f < DFT_NF; ++f) {
            in_r[f] = (smp_k >= 0 && f < dft->nf) ? dft->r_in[f] : real(0);
            in_i[f] = (smp_k >= 0 && f < dft->nf) ? dft->i_in[f] : real(0);
        }
    }

    if constexpr (!EDGE && !FLUX && !DFT && sizeof(real) == 4) {
        // interior warps of the FDTD form in packed arithmetic: the same operations in the same order, two cells per
        // instruction except the two shifted-neighbour differences
        const float2 nz = make_float2(__uint_as_float((unsigned)p.negzero2), __uint_as_float((unsigned)(p.negzero2 >> 32)));
        const float2 half2 = make_float2(0.5f, 0.5f);
        for (int s = 0; s < p.T; ++s) {
            const float hy_left = __shfl_up_sync(FULL, hy[KC - 1], 1);
#pragma unroll
            for (int k = 0; k < KC; k += 2) {
                const float2 curl = make_float2((k == 0 ? hy_left : hy[k == 0 ? 0 : k - 1]) - hy[k], hy[k] - hy[k + 1]);
                const float2 e = pk_add(pk_mul(make_float2(c0[k], c0[k + 1]), make_float2(ex[k], ex[k + 1]), nz),
                                        pk_mul(make_float2(c1[k], c1[k + 1]), curl, nz));
                ex[k] = e.x; ex[k + 1] = e.y;
            }
            if (src_k >= 0 && p.src_field == 0) {
#pragma unroll
                for (int k = 0; k < KC; ++k)
                    if (k == src_k) ex[k] = inject<real>(ex[k], p.src[s], p.src_hard);
            }
            const float ex_right = __shfl_down_sync(FULL, ex[0], 1);
#pragma unroll
            for (int k = 0; k < KC; k += 2) {
                const float2 d = make_float2(ex[k] - ex[k + 1], ex[k + 1] - (k + 2 < KC ? ex[k + 2 < KC ? k + 2 : k] : ex_right));
                const float2 h = pk_add(make_float2(hy[k], hy[k + 1]), pk_mul(half2, d, nz));
                hy[k] = h.x; hy[k + 1] = h.y;
            }
        }
    } else
    for (int s = 0; s < p.T; ++s) {
        // ---- E half step: hy[i-1] of a lane's first cell comes from the lane below
        const real hy_left = __shfl_up_sync(FULL, hy[KC - 1], 1);
        real curl[KC];
#pragma unroll
        for (int k = 0; k < KC; ++k) curl[k] = ((k == 0) ? hy_left : hy[k == 0 ? 0 : k - 1]) - hy[k];
        if (FLUX) {
#pragma unroll
            for (int k = 0; k < KC; ++k)
                if (!EDGE || ((g0 + k >= 1) && (g0 + k < p.nx))) dx[k] = dx[k] + half * curl[k];
            if (src_k >= 0 && p.src_field == 1) {
#pragma unroll
                for (int k = 0; k < KC; ++k)
                    if (k == src_k) dx[k] = inject<real>(dx[k], p.src[s], p.src_hard);
            }
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                if (!EDGE || ((g0 + k >= 1) && (g0 + k < p.nx))) {
                    if (DEBYE) {
                        const real cs = c2[k] * sx[k];
                        ex[k] = c0[k] * ((dx[k] - ix[k]) - cs);
                        sx[k] = cs + (c3[k] * ex[k]);
                    } else {
                        ex[k] = c0[k] * (dx[k] - ix[k]);
                    }
                    ix[k] = ix[k] + c1[k] * ex[k];
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < KC; ++k)
                if (!EDGE || ((g0 + k >= 1) && (g0 + k < p.nx))) ex[k] = (c0[k] * ex[k]) + (c1[k] * curl[k]);
            if (src_k >= 0 && p.src_field == 0) {
#pragma unroll
                for (int k = 0; k < KC; ++k)
                    if (k == src_k) ex[k] = inject<real>(ex[k], p.src[s], p.src_hard);
            }
        }
        // ---- running DFT of the finished E update, BEFORE the ABC (fd1d_2_2.py:138-142: dxfield, exfield, fourier,
        //      hyfield); float64 product and sum, one rounding into the array type per step (= k_fourier)
        if constexpr (DFT) {
#pragma unroll
            for (int f = 0; f < DFT_NF; ++f) {
                if (f < dft->nf) {
                    const double cf = dft->c[f][s], sf = dft->s[f][s];
#pragma unroll
                    for (int k = 0; k < KC; ++k) {
                        const double e = static_cast<double>(ex[k]);
                        acc_r[f][k] = static_cast<real>(static_cast<double>(acc_r[f][k]) + cf * e);
                        acc_i[f][k] = static_cast<real>(static_cast<double>(acc_i[f][k]) - sf * e);
                        if (k == smp_k) {
                            in_r[f] = static_cast<real>(static_cast<double>(in_r[f]) + cf * e);
                            in_i[f] = static_cast<real>(static_cast<double>(in_i[f]) - sf * e);
                        }
                    }
                }
            }
        }
        // ---- two-step-delay ABC on the finished E half step (fd1d_1_2.py:47-48: the right-hand sides are read first)
        if (EDGE && (has_left || has_right)) {
            const real e_above = __shfl_down_sync(FULL, ex[0], 1);          // ex of the cell after my last one
            const real e_below = __shfl_up_sync(FULL, ex[KC - 1], 1);       // ex of the cell before my first one
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                if (k == k_first) {
                    const real e1 = (k == KC - 1) ? e_above : ex[k == KC - 1 ? k : k + 1];
                    ex[k] = b0; b0 = b1; b1 = e1;
                }
                if (k == k_last) {
                    const real e2 = (k == 0) ? e_below : ex[k == 0 ? 0 : k - 1];
                    ex[k] = b3; b3 = b2; b2 = e2;
                }
            }
        }
        // ---- H half step: ex[i+1] of a lane's last cell comes from the lane above
        const real ex_right = __shfl_down_sync(FULL, ex[0], 1);
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            const real er = (k == KC - 1) ? ex_right : ex[k == KC - 1 ? k : k + 1];
            if (!EDGE || ((g0 + k >= 0) && (g0 + k < p.nx - 1))) hy[k] = hy[k] + half * (ex[k] - er);
        }
    }

    // ---- write the owned segment to the other array set (whole vectors: halo and useful are multiples of VEC)
    auto st_cells = [&](real *dst, const real (&d)[KC]) {
        if (!EDGE) {
#pragma unroll
            for (int v = 0; v < KC / VEC; ++v) {
                const int g = g0 + v * VEC;
                if (g >= seg_lo && g + VEC <= seg_hi) {
                    float4 t;
                    real *q = reinterpret_cast<real *>(&t);
#pragma unroll
                    for (int e = 0; e < VEC; ++e) q[e] = d[v * VEC + e];
                    *reinterpret_cast<float4 *>(dst + g) = t;
                }
            }
        } else {
#pragma unroll
            for (int k = 0; k < KC; ++k)
                if (g0 + k >= seg_lo && g0 + k < seg_hi) dst[g0 + k] = d[k];
        }
    };
    st_cells(p.out[0], ex);
    st_cells(p.out[1], hy);
    if (FLUX) {
        st_cells(p.out[2], dx);
        st_cells(p.out[3], ix);
        if (DEBYE) st_cells(p.out[4], sx);
    }
    if constexpr (DFT) {
#pragma unroll
        for (int f = 0; f < DFT_NF; ++f) {
            if (f < dft->nf) {
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    const int g = g0 + k;
                    if (g >= seg_lo && g < seg_hi) {
                        dft->r_pt[(size_t)f * p.nx + g] = acc_r[f][k];
                        dft->i_pt[(size_t)f * p.nx + g] = acc_i[f][k];
                    }
                }
                if (smp_k >= 0) { dft->r_in[f] = in_r[f]; dft->i_in[f] = in_i[f]; }
            }
        }
    }
    if (EDGE && p.abc) {
        // bc[0..1] belong to the warp owning cell 0, bc[2..3] to the one owning cell nx-1
        if (seg_lo == 0 && k_first >= 0) { p.bc_out[0] = b0; p.bc_out[1] = b1; }
        if (seg_hi == p.nx && k_last >= 0) { p.bc_out[2] = b2; p.bc_out[3] = b3; }
    }
}

template <typename real, bool FLUX, bool DEBYE>
__global__ void __launch_bounds__(128) k1_advance(const __grid_constant__ LineParams<real> p) {
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (w >= p.nwarps) return;
    const int base = w * p.useful - p.halo;
    // warp-uniform: whole staged range strictly inside the line (so no end effects) -- vector accesses are then aligned
    const bool edge = (base < 1) || (base + LineShape<real>::W > p.nx - 1);
    if (edge) line_body<real, FLUX, DEBYE, true>(p, w, lane);
    else      line_body<real, FLUX, DEBYE, false>(p, w, lane);
}

// the same pass carrying the running DFT (half the cells per lane, accumulators in registers)
template <typename real, bool FLUX, bool DEBYE>
__global__ void __launch_bounds__(128) k1_advance_dft(const __grid_constant__ LineParams<real> p,
                                                      const __grid_constant__ LineDft<real> d) {
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (w >= p.nwarps) return;
    const int base = w * p.useful - p.halo;
    const bool edge = (base < 1) || (base + LineShape<real, true>::W > p.nx - 1);
    if (edge) line_body<real, FLUX, DEBYE, true, true>(p, w, lane, &d);
    else      line_body<real, FLUX, DEBYE, false, true>(p, w, lane, &d);
}

template <typename real, bool DFT>
int advance1d(const fdtd1d_problem *q, int cur, int nsteps, const double *src, int tblock, cudaStream_t st,
              int *cur_out) {
    using Shape = LineShape<real, DFT>;
    const bool flux = (q->flags & (FDTD_FLUX | FDTD_DEBYE)) != 0, debye = (q->flags & FDTD_DEBYE) != 0;
    const bool has_src = q->src_index >= 0 && src != nullptr;
    // vector accesses need 16-byte aligned arrays; a warp's first staged cell is a whole number of vectors into the line
    int done = 0;
    while (done < nsteps) {
        int T = min(min(tblock, Shape::TMAX), nsteps - done);
        {
            // The right-hand ABC reads ex[nx-2] after every sub-step.  When the segments leave the last warp exactly one
            // cell (nx-1 a multiple of the segment length), ex[nx-2] is that warp's innermost halo cell, valid for
            // halo-1 sub-steps only -- one fewer than the pass takes if T is a whole number of vectors.  Take one step
            // less in this pass (the halo and the segment length stay as they are).
            const int halo_T = ((T + Shape::VEC - 1) / Shape::VEC) * Shape::VEC, use_T = Shape::W - 2 * halo_T;
            if ((q->flags & FDTD_ABC) && T > 1 && halo_T == T && q->nx > use_T && (q->nx - 1) % use_T == 0) --T;
        }
        LineParams<real> lp;
        for (int f = 0; f < 5; ++f) {
            lp.in[f] = (const real *)q->state[cur][f];
            lp.out[f] = (real *)q->state[cur ^ 1][f];
        }
        lp.bc_in = (const real *)q->bc[cur];
        lp.bc_out = (real *)q->bc[cur ^ 1];
        lp.ca = (const real *)q->ca; lp.cb = (const real *)q->cb;
        lp.nax = (const real *)q->md.nax; lp.nbx = (const real *)q->md.nbx;
        lp.ncx = (const real *)q->md.ncx; lp.ndx = (const real *)q->md.ndx;
        lp.nx = q->nx; lp.T = T;
        lp.halo = ((T + Shape::VEC - 1) / Shape::VEC) * Shape::VEC;
        lp.useful = Shape::W - 2 * lp.halo;
        lp.nwarps = (q->nx + lp.useful - 1) / lp.useful;
        lp.abc = (q->flags & FDTD_ABC) != 0;
        lp.src_field = q->src_field; lp.src_index = has_src ? q->src_index : -1; lp.src_hard = q->src_hard;
        for (int k = 0; k < T1MAX; ++k) lp.src[k] = (has_src && k < T) ? src[done + k] : 0.0;
        lp.negzero2 = 0x8000000080000000ull;
        // short lines: one warp per CTA spreads the few warps over the SMs; long lines: 4 warps per CTA
        const int wpc = lp.nwarps >= 8 * fdtd::sm_count() ? 4 : 1;
        const int grid = (lp.nwarps + wpc - 1) / wpc;
        if constexpr (DFT) {
            static_assert(Shape::TMAX <= DFT_TMAX, "phase table too small for the deepest DFT pass");
            LineDft<real> ld;
            ld.r_pt = (real *)q->ft.r_pt; ld.i_pt = (real *)q->ft.i_pt;
            ld.r_in = (q->ft.r_in && q->ft.i_in) ? (real *)q->ft.r_in : nullptr;
            ld.i_in = ld.r_in ? (real *)q->ft.i_in : nullptr;
            ld.nf = q->nf; ld.sample = q->dft_sample;
            for (int f = 0; f < DFT_NF; ++f)
                for (int k = 0; k < DFT_TMAX; ++k) {
                    const bool on = f < q->nf && k < T;
                    ld.c[f][k] = on ? q->dft_cos[(size_t)(done + k) * q->nf + f] : 0.0;
                    ld.s[f][k] = on ? q->dft_sin[(size_t)(done + k) * q->nf + f] : 0.0;
                }
            if (debye)     k1_advance_dft<real, true, true><<<grid, 32 * wpc, 0, st>>>(lp, ld);
            else if (flux) k1_advance_dft<real, true, false><<<grid, 32 * wpc, 0, st>>>(lp, ld);
            else           k1_advance_dft<real, false, false><<<grid, 32 * wpc, 0, st>>>(lp, ld);
            FDTD_LAUNCH_CHECK("k1_advance_dft");
        } else {
            if (debye)     k1_advance<real, true, true><<<grid, 32 * wpc, 0, st>>>(lp);
            else if (flux) k1_advance<real, true, false><<<grid, 32 * wpc, 0, st>>>(lp);
            else           k1_advance<real, false, false><<<grid, 32 * wpc, 0, st>>>(lp);
            FDTD_LAUNCH_CHECK("k1_advance");
        }
        cur ^= 1;
        done += T;
    }
    *cur_out = cur;
    return FDTD_OK;
}

}  // namespace

#define DISPATCH1(dtype, CALL)                                                    \
    if ((dtype) == FDTD_F32) { using real = float; CALL; }                        \
    else if ((dtype) == FDTD_F64) { using real = double; CALL; }                  \
    else { fdtd::set_error("unknown dtype %d", (int)(dtype)); return FDTD_EINVAL; }

extern "C" {

int fdtd1d_exfield(int dtype, int nx, const void *ca, const void *cb, void *ex, const void *hy, const fdtd_source *src,
                   void *stream) {
    FDTD_REQUIRE(nx >= 2 && ex && hy, "fdtd1d_exfield: bad arguments (nx=%d)", nx);
    const bool on = src && src->target;
    FDTD_REQUIRE(!on || (src->target == ex && src->index >= 0 && src->index < nx), "fdtd1d_exfield: source must target ex[0..nx)");
    DISPATCH1(dtype, (k1_exfield<real><<<(nx + 255) / 256, 256, 0, fdtd::as_stream(stream)>>>(
                         nx, (const real *)ca, (const real *)cb, (real *)ex, (const real *)hy, on ? (int)src->index : -1,
                         on ? src->hard : 0, on ? src->value : 0.0)));
    FDTD_LAUNCH_CHECK("k1_exfield");
    return FDTD_OK;
}

int fdtd1d_dxfield(int dtype, int nx, void *dx, const void *hy, const fdtd_source *src, void *stream) {
    FDTD_REQUIRE(nx >= 2 && dx && hy, "fdtd1d_dxfield: bad arguments (nx=%d)", nx);
    const bool on = src && src->target;
    FDTD_REQUIRE(!on || (src->target == dx && src->index >= 0 && src->index < nx), "fdtd1d_dxfield: source must target dx[0..nx)");
    DISPATCH1(dtype, (k1_dxfield<real><<<(nx + 255) / 256, 256, 0, fdtd::as_stream(stream)>>>(
                         nx, (real *)dx, (const real *)hy, on ? (int)src->index : -1, on ? src->hard : 0,
                         on ? src->value : 0.0)));
    FDTD_LAUNCH_CHECK("k1_dxfield");
    return FDTD_OK;
}

int fdtd1d_exfield_flux(int dtype, int nx, const fdtd_medium1d *md, const void *dx, void *ix, void *sx, void *ex,
                        void *stream) {
    FDTD_REQUIRE(nx >= 2 && md && md->nax && md->nbx && dx && ix && ex, "fdtd1d_exfield_flux: bad arguments");
    FDTD_REQUIRE(!sx || (md->ncx && md->ndx), "fdtd1d_exfield_flux: sx given without ncx/ndx");
    DISPATCH1(dtype, (k1_exfield_flux<real><<<(nx + 255) / 256, 256, 0, fdtd::as_stream(stream)>>>(
                         nx, (const real *)md->nax, (const real *)md->nbx, (const real *)md->ncx, (const real *)md->ndx,
                         (const real *)dx, (real *)ix, (real *)sx, (real *)ex)));
    FDTD_LAUNCH_CHECK("k1_exfield_flux");
    return FDTD_OK;
}

int fdtd1d_hyfield(int dtype, int nx, void *ex, void *hy, void *bc, int abc, void *stream) {
    FDTD_REQUIRE(nx >= 3 && ex && hy && (!abc || bc), "fdtd1d_hyfield: bad arguments (nx=%d)", nx);
    cudaStream_t st = fdtd::as_stream(stream);
    if (abc) {
        DISPATCH1(dtype, (k1_abc<real><<<1, 1, 0, st>>>(nx, (real *)ex, (real *)bc)));
        FDTD_LAUNCH_CHECK("k1_abc");
    }
    DISPATCH1(dtype, (k1_hy<real><<<(nx + 255) / 256, 256, 0, st>>>(nx, (const real *)ex, (real *)hy)));
    FDTD_LAUNCH_CHECK("k1_hy");
    return FDTD_OK;
}

int fdtd1d_fourier(int dtype, int nf, int nx, const double *cosv, const double *sinv, const void *ex, int sample_index,
                   const fdtd_ftrans *ft, void *stream) {
    FDTD_REQUIRE(nx >= 1 && sample_index >= 0 && sample_index < nx, "fdtd1d_fourier: bad nx=%d / sample index %d", nx, sample_index);
    const size_t esz = dtype == FDTD_F64 ? 8 : 4;
    const void *sample = ex ? (const char *)ex + esz * (size_t)sample_index : nullptr;
    return fdtd::launch_fourier(dtype, nf, (size_t)nx, cosv, sinv, ex, sample, ft, fdtd::as_stream(stream));
}

int fdtd1d_advance(const fdtd1d_problem *q, int cur, int nsteps, const double *src, int tblock, void *stream,
                   int *cur_out) {
    FDTD_REQUIRE(q && cur_out, "fdtd1d_advance: null problem / cur_out");
    FDTD_REQUIRE(cur == 0 || cur == 1, "fdtd1d_advance: cur must be 0 or 1");
    FDTD_REQUIRE(q->nx >= 3, "fdtd1d_advance: nx=%d too small", q->nx);
    FDTD_REQUIRE(tblock >= 1 && tblock <= T1MAX, "fdtd1d_advance: tblock %d outside [1, %d]", tblock, T1MAX);
    FDTD_REQUIRE(nsteps >= 0, "fdtd1d_advance: nsteps < 0");
    const bool flux = (q->flags & (FDTD_FLUX | FDTD_DEBYE)) != 0, debye = (q->flags & FDTD_DEBYE) != 0;
    const int nfields = debye ? 5 : (flux ? 4 : 2);
    for (int s = 0; s < 2; ++s) {
        for (int f = 0; f < nfields; ++f)
            FDTD_REQUIRE(q->state[s][f] && fdtd::aligned16(q->state[s][f]), "fdtd1d_advance: state[%d][%d] null or not 16-byte aligned", s, f);
        FDTD_REQUIRE(!(q->flags & FDTD_ABC) || q->bc[s], "fdtd1d_advance: FDTD_ABC needs bc[%d]", s);
    }
    FDTD_REQUIRE(!flux || (q->md.nax && q->md.nbx), "fdtd1d_advance: flux form needs nax/nbx");
    {
        const void *coef[6] = {q->ca, q->cb, q->md.nax, q->md.nbx, q->md.ncx, q->md.ndx};
        for (int k = 0; k < 6; ++k) FDTD_REQUIRE(fdtd::aligned16(coef[k]), "fdtd1d_advance: coefficient array %d not 16-byte aligned", k);
    }
    FDTD_REQUIRE(!debye || (q->md.ncx && q->md.ndx), "fdtd1d_advance: Debye form needs ncx/ndx");
    FDTD_REQUIRE(q->src_index < q->nx, "fdtd1d_advance: source index %d outside the line", q->src_index);
    FDTD_REQUIRE(q->src_field == 0 || (q->src_field == 1 && flux), "fdtd1d_advance: src_field %d invalid for this form", q->src_field);
    FDTD_REQUIRE(q->nf >= 0 && q->nf <= DFT_NF, "fdtd1d_advance: nf=%d outside [0, %d] (more frequencies: fdtd1d_fourier after every step)", q->nf, DFT_NF);
    if (q->nf > 0) {
        FDTD_REQUIRE(q->ft.r_pt && q->ft.i_pt && q->dft_cos && q->dft_sin, "fdtd1d_advance: running DFT needs ft.r_pt / ft.i_pt and the phase tables");
        FDTD_REQUIRE(!(q->ft.r_in && q->ft.i_in) || (q->dft_sample >= 0 && q->dft_sample < q->nx),
                     "fdtd1d_advance: DFT sample cell %d outside the line", q->dft_sample);
        if (q->dtype == FDTD_F32) return advance1d<float, true>(q, cur, nsteps, src, tblock, fdtd::as_stream(stream), cur_out);
        if (q->dtype == FDTD_F64) return advance1d<double, true>(q, cur, nsteps, src, tblock, fdtd::as_stream(stream), cur_out);
    }
    if (q->dtype == FDTD_F32) return advance1d<float, false>(q, cur, nsteps, src, tblock, fdtd::as_stream(stream), cur_out);
    if (q->dtype == FDTD_F64) return advance1d<double, false>(q, cur, nsteps, src, tblock, fdtd::as_stream(stream), cur_out);
    fdtd::set_error("fdtd1d_advance: unknown dtype %d", q->dtype);
    return FDTD_EINVAL;
}

}  // extern "C"
