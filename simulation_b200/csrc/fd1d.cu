// 1D Ex/Hy line: the reference-named step functions (one kernel each) and a fused, temporally blocked
// advance.  Update rules, ranges and evaluation order follow the reference numpy programs
// (fd1d/program/fd1d_1_5.py:63-72 FDTD form, fd1d_2_3.py:73-94 flux/Debye form); no FMA contraction.
//
// Fused path: every CTA owns SEG cells, loads them plus T halo cells per side into shared memory (Ex, Hy)
// and registers (pointwise state and coefficients), runs T leap-frog steps on chip and writes its SEG cells
// to the other (ping-pong) array set -- 2*T-cell recompute per CTA instead of a grid-wide barrier per step.
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
constexpr int NT = 256;       // threads per CTA
constexpr int KC = 8;         // cells per thread  -> NT*KC cells staged per CTA
constexpr int T1MAX = 64;     // deepest time block

template <typename real>
struct LineParams {
    const real *in[5];        // ex, hy, dx, ix, sx
    real *out[5];
    const real *bc_in;
    real *bc_out;
    const real *ca, *cb, *nax, *nbx, *ncx, *ndx;
    int nx, T, seg;           // seg = cells produced per CTA = NT*KC - 2*T
    int abc, src_field, src_index, src_hard;
    double src[T1MAX];        // waveform samples of this pass (by value: no staging buffer to manage)
};

// The T leap-frog steps of one CTA.  Every thread keeps the Ex / Hy of its own KC cells (slot c = k*NT + tid) in
// registers next to the pointwise state; shared memory only carries values to the neighbouring cell
// (hy[c-1] for the E half step, ex[c+1] for the H half step): 2 LDS + 2 STS per cell-step.  Both shared arrays are
// shifted by one slot so that c-1 / c+1 of the first / last staged cell stay in bounds (halo garbage, never used).
// EDGE: the staged range touches an end of the line (cells outside the line, the never-updated ex[0], ex[nx-1],
// hy[nx-1], the ABC); interior CTAs run without any mask.
template <typename real, bool FLUX, bool DEBYE, bool EDGE>
__device__ __forceinline__ void line_body(const LineParams<real> &p, real *s_ex, real *s_hy, real *s_bc) {
    const int tid = threadIdx.x;
    const int seg_lo = blockIdx.x * p.seg;
    const int seg_hi = min(seg_lo + p.seg, p.nx);
    const int base = seg_lo - p.T;                 // global cell of staged slot 0 (may be negative)
    const real half = real(0.5);

    real ex[KC], hy[KC], dx[KC], ix[KC], sx[KC], c0[KC], c1[KC], c2[KC], c3[KC];
    int src_k = -1;                                // which of this thread's cells carries the source (-1: none)
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        const int c = k * NT + tid, g = base + c;
        const bool in = !EDGE || ((g >= 0) && (g < p.nx));
        ex[k] = in ? p.in[0][g] : real(0);
        hy[k] = in ? p.in[1][g] : real(0);
        dx[k] = ix[k] = sx[k] = c2[k] = c3[k] = real(0);
        if (FLUX) {
            dx[k] = in ? p.in[2][g] : real(0);
            ix[k] = in ? p.in[3][g] : real(0);
            c0[k] = in ? p.nax[g] : real(1);
            c1[k] = in ? p.nbx[g] : real(0);
            if (DEBYE) {
                sx[k] = in ? p.in[4][g] : real(0);
                c2[k] = in ? p.ncx[g] : real(0);
                c3[k] = in ? p.ndx[g] : real(0);
            }
        } else {
            c0[k] = (in && p.ca) ? p.ca[g] : real(1);
            c1[k] = (in && p.cb) ? p.cb[g] : real(0.5);
        }
        if (p.src_index >= 0 && g == p.src_index) src_k = k;
        s_hy[c + 1] = hy[k];
        s_ex[c + 1] = ex[k];
    }
    if (tid == 0) { s_hy[0] = s_ex[0] = real(0); s_hy[NT * KC + 1] = s_ex[NT * KC + 1] = real(0); }
    if (tid < 4) s_bc[tid] = p.abc ? p.bc_in[tid] : real(0);
    __syncthreads();

    const bool has_left = EDGE && p.abc && (base <= 0);                       // staged range contains cells 0, 1
    const bool has_right = EDGE && p.abc && (base + NT * KC >= p.nx);         // ... and cells nx-2, nx-1
    for (int s = 0; s < p.T; ++s) {
        // ---- E half step: ex (or dx -> ex) of the own cells; hy[i-1] comes from the neighbour through smem
        real curl[KC];
#pragma unroll
        for (int k = 0; k < KC; ++k) curl[k] = s_hy[k * NT + tid] - hy[k];     // slot c-1 lives at index c
        if (FLUX) {
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                const int g = base + k * NT + tid;
                if (!EDGE || ((g >= 1) && (g < p.nx))) dx[k] = dx[k] + half * curl[k];
            }
            if (src_k >= 0 && p.src_field == 1) {          // at most one thread of the CTA
#pragma unroll
                for (int k = 0; k < KC; ++k)
                    if (k == src_k) dx[k] = inject<real>(dx[k], p.src[s], p.src_hard);
            }
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                const int g = base + k * NT + tid;
                if (!EDGE || ((g >= 1) && (g < p.nx))) {
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
            for (int k = 0; k < KC; ++k) {
                const int g = base + k * NT + tid;
                if (!EDGE || ((g >= 1) && (g < p.nx))) ex[k] = (c0[k] * ex[k]) + (c1[k] * curl[k]);
            }
            if (src_k >= 0 && p.src_field == 0) {
#pragma unroll
                for (int k = 0; k < KC; ++k)
                    if (k == src_k) ex[k] = inject<real>(ex[k], p.src[s], p.src_hard);
            }
        }
#pragma unroll
        for (int k = 0; k < KC; ++k) s_ex[k * NT + tid + 1] = ex[k];
        __syncthreads();
        // ---- two-step-delay ABC (one thread; needs the finished E half step), then its owners re-read
        if (EDGE && (has_left || has_right)) {
            if (tid == 0) {
                if (has_left) {
                    const int o = -base + 1;                                   // index of cell 0
                    const real e1 = s_ex[o + 1], b0 = s_bc[0], b1 = s_bc[1];
                    s_ex[o] = b0; s_bc[0] = b1; s_bc[1] = e1;
                }
                if (has_right) {
                    const int o = p.nx - 1 - base + 1;                         // index of cell nx-1
                    const real e2 = s_ex[o - 1], b3 = s_bc[3], b2 = s_bc[2];
                    s_ex[o] = b3; s_bc[3] = b2; s_bc[2] = e2;
                }
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < KC; ++k) {
                const int g = base + k * NT + tid;
                if (g == 0 || g == p.nx - 1) ex[k] = s_ex[k * NT + tid + 1];
            }
        }
        // ---- H half step: ex[i+1] comes from the neighbour through smem
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            const int g = base + k * NT + tid;
            const real er = s_ex[k * NT + tid + 2];                            // slot c+1 lives at index c+2
            if (!EDGE || ((g >= 0) && (g < p.nx - 1))) hy[k] = hy[k] + half * (ex[k] - er);
        }
#pragma unroll
        for (int k = 0; k < KC; ++k) s_hy[k * NT + tid + 1] = hy[k];
        __syncthreads();
    }

    // ---- write the owned segment to the other array set
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        const int g = base + k * NT + tid;
        if (g >= seg_lo && g < seg_hi) {
            p.out[0][g] = ex[k];
            p.out[1][g] = hy[k];
            if (FLUX) {
                p.out[2][g] = dx[k];
                p.out[3][g] = ix[k];
                if (DEBYE) p.out[4][g] = sx[k];
            }
        }
    }
    if (p.abc && tid < 4) {
        // bc[0..1] belong to the CTA owning cell 0, bc[2..3] to the one owning cell nx-1
        const bool mine = (tid < 2) ? (seg_lo == 0) : (seg_hi == p.nx);
        if (mine) p.bc_out[tid] = s_bc[tid];
    }
}

template <typename real, bool FLUX, bool DEBYE>
__global__ void __launch_bounds__(NT) k1_advance(const __grid_constant__ LineParams<real> p) {
    __shared__ real s_ex[NT * KC + 2];
    __shared__ real s_hy[NT * KC + 2];
    __shared__ real s_bc[4];
    const int base = blockIdx.x * p.seg - p.T;
    const bool edge = (base < 1) || (base + NT * KC > p.nx - 1);      // CTA-uniform
    if (edge) line_body<real, FLUX, DEBYE, true>(p, s_ex, s_hy, s_bc);
    else      line_body<real, FLUX, DEBYE, false>(p, s_ex, s_hy, s_bc);
}

template <typename real>
int advance1d(const fdtd1d_problem *q, int cur, int nsteps, const double *src, int tblock, cudaStream_t st,
              int *cur_out) {
    const bool flux = (q->flags & (FDTD_FLUX | FDTD_DEBYE)) != 0, debye = (q->flags & FDTD_DEBYE) != 0;
    const bool has_src = q->src_index >= 0 && src != nullptr;
    int done = 0;
    while (done < nsteps) {
        const int T = min(tblock, nsteps - done);
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
        lp.nx = q->nx; lp.T = T; lp.seg = NT * KC - 2 * T;
        lp.abc = (q->flags & FDTD_ABC) != 0;
        lp.src_field = q->src_field; lp.src_index = has_src ? q->src_index : -1; lp.src_hard = q->src_hard;
        for (int k = 0; k < T1MAX; ++k) lp.src[k] = (has_src && k < T) ? src[done + k] : 0.0;
        const int grid = (q->nx + lp.seg - 1) / lp.seg;
        if (debye)     k1_advance<real, true, true><<<grid, NT, 0, st>>>(lp);
        else if (flux) k1_advance<real, true, false><<<grid, NT, 0, st>>>(lp);
        else           k1_advance<real, false, false><<<grid, NT, 0, st>>>(lp);
        FDTD_LAUNCH_CHECK("k1_advance");
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
        for (int f = 0; f < nfields; ++f) FDTD_REQUIRE(q->state[s][f], "fdtd1d_advance: state[%d][%d] is null", s, f);
        FDTD_REQUIRE(!(q->flags & FDTD_ABC) || q->bc[s], "fdtd1d_advance: FDTD_ABC needs bc[%d]", s);
    }
    FDTD_REQUIRE(!flux || (q->md.nax && q->md.nbx), "fdtd1d_advance: flux form needs nax/nbx");
    FDTD_REQUIRE(!debye || (q->md.ncx && q->md.ndx), "fdtd1d_advance: Debye form needs ncx/ndx");
    FDTD_REQUIRE(q->src_index < q->nx, "fdtd1d_advance: source index %d outside the line", q->src_index);
    FDTD_REQUIRE(q->src_field == 0 || (q->src_field == 1 && flux), "fdtd1d_advance: src_field %d invalid for this form", q->src_field);
    if (q->dtype == FDTD_F32) return advance1d<float>(q, cur, nsteps, src, tblock, fdtd::as_stream(stream), cur_out);
    if (q->dtype == FDTD_F64) return advance1d<double>(q, cur, nsteps, src, tblock, fdtd::as_stream(stream), cur_out);
    fdtd::set_error("fdtd1d_advance: unknown dtype %d", q->dtype);
    return FDTD_EINVAL;
}

}  // extern "C"
