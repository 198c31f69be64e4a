// 2D TM step functions under the reference's names, one kernel per reference function, on raw device
// pointers.  This is the drop-in / parity / debug path (the benchmarked path is fd2d_march.cu); the
// arithmetic, index ranges and evaluation order are those of the reference numpy programs
// (fd2d/program/fd2d_3_3.py:60-110, fd2d/python/fd2d_3_4.py:131-136), so the results are bit-identical.
#include "common.cuh"

namespace {

using fdtd::inject;

constexpr int BX = 64, BY = 4;   // 64 consecutive j per warp-pair row: coalesced along the fast axis

template <typename real>
struct Pml {
    const real *fx1, *fx2, *fx3, *fy1, *fy2, *fy3, *gx2, *gx3, *gy2, *gy3;
    explicit Pml(const fdtd_pmlayer &p)
        : fx1((const real *)p.fx1), fx2((const real *)p.fx2), fx3((const real *)p.fx3),
          fy1((const real *)p.fy1), fy2((const real *)p.fy2), fy3((const real *)p.fy3),
          gx2((const real *)p.gx2), gx3((const real *)p.gx3), gy2((const real *)p.gy2),
          gy3((const real *)p.gy3) {}
};

template <typename real>
__global__ void k_source(real *target, long long index, int hard, double value) {
    target[index] = inject<real>(target[index], value, hard);
}

// ezi[1:ny] += 0.5*(hxi[j-1]-hxi[j]); then the two-step-delay ABC on both ends.  Single CTA: the ABC
// reads ezi[1] / ezi[ny-2] AFTER the stencil, which needs a barrier (the reference CUDA kernel races here).
template <typename real>
__global__ void k_ezinct(int ny, real *ezi, const real *hxi, real *bc) {
    for (int j = 1 + threadIdx.x; j < ny; j += blockDim.x) ezi[j] = ezi[j] + real(0.5) * (hxi[j - 1] - hxi[j]);
    __syncthreads();
    if (threadIdx.x == 0) {
        real e1 = ezi[1], b0 = bc[0], b1 = bc[1];
        ezi[0] = b0; bc[0] = b1; bc[1] = e1;
        real e2 = ezi[ny - 2], b3 = bc[3], b2 = bc[2];
        ezi[ny - 1] = b3; bc[3] = b2; bc[2] = e2;
    }
}

template <typename real>
__global__ void k_hxinct(int ny, const real *ezi, real *hxi) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < ny - 1) hxi[j] = hxi[j] + real(0.5) * (ezi[j] - ezi[j + 1]);
}

template <typename real>
__global__ void k_dfield(int nx, int ny, Pml<real> p, real *dz, const real *hx, const real *hy) {
    int j = blockIdx.x * BX + threadIdx.x;
    int i = blockIdx.y * BY + threadIdx.y;
    if (i < 1 || i >= nx || j < 1 || j >= ny) return;
    size_t n = (size_t)i * ny + j;
    real curl = ((hy[n] - hy[n - ny]) - hx[n]) + hx[n - 1];
    dz[n] = ((p.gx3[i] * p.gy3[j]) * dz[n]) + (((p.gx2[i] * p.gy2[j]) * real(0.5)) * curl);
}

template <typename real>
__global__ void k_inctdz(int nx, int ny, int npml, const real *hxi, real *dz) {
    int i = npml - 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nx - npml) return;
    size_t a = (size_t)i * ny + (npml - 1), b = (size_t)i * ny + (ny - npml);
    dz[a] = dz[a] + real(0.5) * hxi[npml - 2];
    dz[b] = dz[b] - real(0.5) * hxi[ny - npml];
}

template <typename real, bool LOSSY>
__global__ void k_efield(int nx, int ny, const real *naz, const real *nbz, const real *dz, real *iz, real *ez) {
    int j = blockIdx.x * BX + threadIdx.x;
    int i = blockIdx.y * BY + threadIdx.y;
    if (i >= nx || j >= ny) return;
    size_t n = (size_t)i * ny + j;
    if (LOSSY) {
        real e = naz[n] * (dz[n] - iz[n]);
        ez[n] = e;
        iz[n] = iz[n] + nbz[n] * e;
    } else {
        ez[n] = naz[n] * dz[n];
    }
}

template <typename real>
__global__ void k_hfield(int nx, int ny, Pml<real> p, const real *ez, real *ihx, real *ihy, real *hx, real *hy) {
    int j = blockIdx.x * BX + threadIdx.x;
    int i = blockIdx.y * BY + threadIdx.y;
    if (i >= nx - 1 || j >= ny - 1) return;
    size_t n = (size_t)i * ny + j;
    real e = ez[n];
    real cm = e - ez[n + 1];
    real cn = e - ez[n + ny];
    real ax = ihx[n] + cm;
    real ay = ihy[n] + cn;
    ihx[n] = ax;
    ihy[n] = ay;
    hx[n] = (p.fy3[j] * hx[n]) + (p.fy2[j] * ((real(0.5) * cm) + (p.fx1[i] * ax)));
    hy[n] = (p.fx3[i] * hy[n]) - (p.fx2[i] * ((real(0.5) * cn) + (p.fy1[j] * ay)));
}

template <typename real>
__global__ void k_incthx(int nx, int ny, int npml, const real *ezi, real *hx) {
    int i = npml - 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i > nx - npml) return;
    size_t a = (size_t)i * ny + (npml - 2), b = (size_t)i * ny + (ny - npml);
    hx[a] = hx[a] + real(0.5) * ezi[npml - 1];
    hx[b] = hx[b] - real(0.5) * ezi[ny - npml];
}

template <typename real>
__global__ void k_incthy(int nx, int ny, int npml, const real *ezi, real *hy) {
    int j = npml - 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (j > ny - npml) return;
    size_t a = (size_t)(npml - 2) * ny + j, b = (size_t)(nx - npml) * ny + j;
    real h = real(0.5) * ezi[j];
    if (a == b) {            // degenerate geometry: both corrections hit the same cell, in reference order
        hy[a] = (hy[a] - h) + h;
    } else {
        hy[a] = hy[a] - h;
        hy[b] = hy[b] + h;
    }
}

// running DFT of Ez at every cell: r_pt[n] += cos_n * ez, i_pt[n] -= sin_n * ez, evaluated in float64 and rounded
// into the array type (what numpy / numba do for float32 arrays: the phase factors are float64).
struct Phases { double c[FDTD_MAX_FREQS], s[FDTD_MAX_FREQS]; };

template <typename real>
__global__ void k_fourier(int nf, size_t ncells, Phases ph, const real *ez, real *r_pt, real *i_pt) {
    const size_t n = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= ncells) return;
    const double e = static_cast<double>(ez[n]);
    for (int f = 0; f < nf; ++f) {
        const size_t m = (size_t)f * ncells + n;
        r_pt[m] = static_cast<real>(static_cast<double>(r_pt[m]) + ph.c[f] * e);
        i_pt[m] = static_cast<real>(static_cast<double>(i_pt[m]) - ph.s[f] * e);
    }
}

template <typename real>
__global__ void k_fourier_source(int nf, Phases ph, const real *sample, real *r_in, real *i_in) {
    const int f = threadIdx.x;
    if (f >= nf) return;
    const double e = static_cast<double>(*sample);
    r_in[f] = static_cast<real>(static_cast<double>(r_in[f]) + ph.c[f] * e);
    i_in[f] = static_cast<real>(static_cast<double>(i_in[f]) - ph.s[f] * e);
}

// Lossy dielectric cylinder, 3x3 sub-cell average, evaluated in float64 exactly as the reference's Python
// statements (fd2d/python/fd2d_3_4.py:173-194; the C/CUDA variants' integer m/3 is a known divergence):
// x = nx/2-1-i+m/3, y = ny/2-1-j+n/3, inside if sqrt(x*x+y*y) <= rgrid.  Rows [row_lo, row_hi) of the global grid.
// ---- pmlparam on the device (fd2d/program/fd2d_3_3.py:113-122): the ten 1D PML vectors from the reference's Python
// statements, evaluated in float64 and rounded on store.  Python's ``x**3`` is the host libm's pow(), which is NOT
// correctly rounded (glibc: < 1 ulp; 50 of the 93 324 layer arguments of npml <= 300 come out one ulp off the exact
// cube), so bit-identity in float64 needs the host's own cubes: the caller may pass them (`cubes`, 2*npml doubles on
// the device: ((npml-n)/npml)**3 for n = 0..npml-1, then ((npml-n-0.5)/npml)**3) and the kernel expands them into the
// ten vectors; without a table the cube is evaluated here in double-double (correctly rounded: float32 vectors equal
// surface.pmlparam bit for bit for every n of every npml <= 300, float64 ones to one ulp of the cube).
__device__ __forceinline__ double cube_rn(const double x) {
    const double hi = x * x, lo = fma(x, x, -hi);            // x^2 = hi + lo exactly
    const double p = hi * x, e = fma(hi, x, -p);             // hi*x = p + e exactly
    return p + (e + lo * x);
}

template <typename real>
__device__ __forceinline__ void pml_entries(const int i, const int N, const int npml, const double *cubes, real *f1,
                                            real *f2, real *f3, real *g2, real *g3) {
    // f-vectors: entries n and N-2-n (H lives on the half cell); g-vectors: n and N-1-n; n = 0 .. npml-1, and a later n
    // of the reference's loop overwrites an earlier one where the two ends meet
    int nf = -1, ng = -1;
    if (i < npml) nf = ng = i;
    const int mf = N - 2 - i, mg = N - 1 - i;
    if (mf >= 0 && mf < npml && mf > nf) nf = mf;
    if (mg >= 0 && mg < npml && mg > ng) ng = mg;
    if (nf >= 0) {
        const double xn = 0.33 * (cubes ? cubes[npml + nf] : cube_rn(((double)(npml - nf) - 0.5) / (double)npml));
        f1[i] = static_cast<real>(xn);
        f2[i] = static_cast<real>(1.0 / (1.0 + xn));
        f3[i] = static_cast<real>((1.0 - xn) / (1.0 + xn));
    } else {
        f1[i] = real(0); f2[i] = real(1); f3[i] = real(1);
    }
    if (ng >= 0) {
        const double xm = 0.33 * (cubes ? cubes[ng] : cube_rn((double)(npml - ng) / (double)npml));
        g2[i] = static_cast<real>(1.0 / (1.0 + xm));
        g3[i] = static_cast<real>((1.0 - xm) / (1.0 + xm));
    } else {
        g2[i] = real(1); g3[i] = real(1);
    }
}

template <typename real>
__global__ void k_pmlparam(int nx, int ny, int npml, const double *cubes, real *fx1, real *fx2, real *fx3, real *fy1, real *fy2, real *fy3,
                           real *gx2, real *gx3, real *gy2, real *gy3) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nx) pml_entries<real>(i, nx, npml, cubes, fx1, fx2, fx3, gx2, gx3);
    if (i < ny) pml_entries<real>(i, ny, npml, cubes, fy1, fy2, fy3, gy2, gy3);
}

template <typename real>
__global__ void k_cylinder(int nx, int ny, int npml, int rgrid, double dt, double epsr, double sigma, int row_lo,
                           int row_hi, real *naz, real *nbz) {
    const int j = blockIdx.x * BX + threadIdx.x;
    const int i = row_lo + blockIdx.y * BY + threadIdx.y;
    if (i >= row_hi || j >= ny) return;
    const size_t n = (size_t)(i - row_lo) * ny + j;
    double vn = 1.0, vb = 0.0;
    if (i >= npml && i < nx - npml && j >= npml && j < ny - npml) {
        const double eps0 = 8.854e-12, de = (epsr - 1) / 9, dc = sigma / 9;
        double epsn = 1.0, cond = 0.0;
        for (int m = -1; m < 2; ++m)
            for (int q = -1; q < 2; ++q) {
                const double x = (((double)nx / 2 - 1) - (double)i) + (double)m / 3;
                const double y = (((double)ny / 2 - 1) - (double)j) + (double)q / 3;
                if (sqrt(x * x + y * y) <= (double)rgrid) { epsn += de; cond += dc; }
            }
        vn = 1 / (epsn + cond * dt / eps0);
        vb = cond * dt / eps0;
    }
    naz[n] = static_cast<real>(vn);
    nbz[n] = static_cast<real>(vb);
}

inline dim3 grid2d(int nx, int ny) { return dim3((ny + BX - 1) / BX, (nx + BY - 1) / BY); }

template <typename real>
int apply_source(const fdtd_source *src, cudaStream_t st) {
    if (src == nullptr || src->target == nullptr) return FDTD_OK;
    FDTD_REQUIRE(src->index >= 0, "source index %lld < 0", src->index);
    k_source<real><<<1, 1, 0, st>>>((real *)src->target, src->index, src->hard, src->value);
    FDTD_LAUNCH_CHECK("k_source");
    return FDTD_OK;
}

bool tfsf_geometry_ok(int nx, int ny, int npml) { return npml >= 2 && 2 * npml <= nx && 2 * npml <= ny; }

}  // namespace

namespace fdtd {
// shared by the 1D and 2D entry points: `ncells` field samples, source sample at sample[0]
int launch_fourier(int dtype, int nf, size_t ncells, const double *cosv, const double *sinv, const void *field,
                   const void *sample, const fdtd_ftrans *ft, cudaStream_t st) {
    FDTD_REQUIRE(nf >= 1 && nf <= FDTD_MAX_FREQS, "fourier: nf=%d outside [1, %d]", nf, FDTD_MAX_FREQS);
    FDTD_REQUIRE(cosv && sinv && field && ft && ft->r_pt && ft->i_pt, "fourier: null argument");
    Phases ph;
    for (int f = 0; f < FDTD_MAX_FREQS; ++f) { ph.c[f] = f < nf ? cosv[f] : 0.0; ph.s[f] = f < nf ? sinv[f] : 0.0; }
    const unsigned blocks = (unsigned)((ncells + 255) / 256);
    if (dtype == FDTD_F32) {
        k_fourier<float><<<blocks, 256, 0, st>>>(nf, ncells, ph, (const float *)field, (float *)ft->r_pt, (float *)ft->i_pt);
        if (sample && ft->r_in && ft->i_in) k_fourier_source<float><<<1, 32, 0, st>>>(nf, ph, (const float *)sample, (float *)ft->r_in, (float *)ft->i_in);
    } else if (dtype == FDTD_F64) {
        k_fourier<double><<<blocks, 256, 0, st>>>(nf, ncells, ph, (const double *)field, (double *)ft->r_pt, (double *)ft->i_pt);
        if (sample && ft->r_in && ft->i_in) k_fourier_source<double><<<1, 32, 0, st>>>(nf, ph, (const double *)sample, (double *)ft->r_in, (double *)ft->i_in);
    } else {
        set_error("fourier: unknown dtype %d", dtype);
        return FDTD_EINVAL;
    }
    FDTD_LAUNCH_CHECK("k_fourier");
    return FDTD_OK;
}

int launch_source(int dtype, const fdtd_source *src, cudaStream_t st) {
    return dtype == FDTD_F32 ? apply_source<float>(src, st) : apply_source<double>(src, st);
}
}  // namespace fdtd

#define DISPATCH(dtype, CALL)                                                     \
    if ((dtype) == FDTD_F32) { using real = float; CALL; }                        \
    else if ((dtype) == FDTD_F64) { using real = double; CALL; }                  \
    else { fdtd::set_error("unknown dtype %d", (int)(dtype)); return FDTD_EINVAL; }

extern "C" {

int fdtd2d_ezinct(int dtype, int ny, void *ezi, const void *hxi, void *bc, void *stream) {
    FDTD_REQUIRE(ny >= 3 && ezi && hxi && bc, "fdtd2d_ezinct: bad arguments (ny=%d)", ny);
    DISPATCH(dtype, (k_ezinct<real><<<1, 1024, 0, fdtd::as_stream(stream)>>>(ny, (real *)ezi, (const real *)hxi, (real *)bc)));
    FDTD_LAUNCH_CHECK("k_ezinct");
    return FDTD_OK;
}

int fdtd2d_hxinct(int dtype, int ny, const void *ezi, void *hxi, void *stream) {
    FDTD_REQUIRE(ny >= 2 && ezi && hxi, "fdtd2d_hxinct: bad arguments (ny=%d)", ny);
    DISPATCH(dtype, (k_hxinct<real><<<(ny + 255) / 256, 256, 0, fdtd::as_stream(stream)>>>(ny, (const real *)ezi, (real *)hxi)));
    FDTD_LAUNCH_CHECK("k_hxinct");
    return FDTD_OK;
}

int fdtd2d_dfield(int dtype, int nx, int ny, const fdtd_pmlayer *pml, void *dz, const void *hx, const void *hy,
                  const fdtd_source *src, void *stream) {
    FDTD_REQUIRE(nx >= 2 && ny >= 2 && pml && dz && hx && hy, "fdtd2d_dfield: bad arguments (nx=%d ny=%d)", nx, ny);
    cudaStream_t st = fdtd::as_stream(stream);
    DISPATCH(dtype, (k_dfield<real><<<grid2d(nx, ny), dim3(BX, BY), 0, st>>>(nx, ny, Pml<real>(*pml), (real *)dz, (const real *)hx, (const real *)hy)));
    FDTD_LAUNCH_CHECK("k_dfield");
    return fdtd::launch_source(dtype, src, st);
}

int fdtd2d_inctdz(int dtype, int nx, int ny, int npml, const void *hxi, void *dz, void *stream) {
    FDTD_REQUIRE(tfsf_geometry_ok(nx, ny, npml) && hxi && dz, "fdtd2d_inctdz: bad geometry nx=%d ny=%d npml=%d", nx, ny, npml);
    int n = nx - 2 * npml + 2;
    DISPATCH(dtype, (k_inctdz<real><<<(n + 255) / 256, 256, 0, fdtd::as_stream(stream)>>>(nx, ny, npml, (const real *)hxi, (real *)dz)));
    FDTD_LAUNCH_CHECK("k_inctdz");
    return FDTD_OK;
}

int fdtd2d_efield(int dtype, int nx, int ny, const fdtd_medium2d *md, const void *dz, void *iz, void *ez, void *stream) {
    FDTD_REQUIRE(nx >= 1 && ny >= 1 && md && md->naz && dz && ez, "fdtd2d_efield: bad arguments");
    FDTD_REQUIRE(iz == nullptr || md->nbz != nullptr, "fdtd2d_efield: iz given without nbz");
    cudaStream_t st = fdtd::as_stream(stream);
    if (iz) {
        DISPATCH(dtype, (k_efield<real, true><<<grid2d(nx, ny), dim3(BX, BY), 0, st>>>(nx, ny, (const real *)md->naz, (const real *)md->nbz, (const real *)dz, (real *)iz, (real *)ez)));
    } else {
        DISPATCH(dtype, (k_efield<real, false><<<grid2d(nx, ny), dim3(BX, BY), 0, st>>>(nx, ny, (const real *)md->naz, nullptr, (const real *)dz, nullptr, (real *)ez)));
    }
    FDTD_LAUNCH_CHECK("k_efield");
    return FDTD_OK;
}

int fdtd2d_fourier(int dtype, int nf, int nx, int ny, const double *cosv, const double *sinv, const void *ezi,
                   int sample_index, const void *ez, const fdtd_ftrans *ft, void *stream) {
    FDTD_REQUIRE(nx >= 1 && ny >= 1, "fdtd2d_fourier: bad grid %dx%d", nx, ny);
    FDTD_REQUIRE(!ezi || (sample_index >= 0 && sample_index < ny), "fdtd2d_fourier: sample index %d outside the incident line", sample_index);
    const size_t esz = dtype == FDTD_F64 ? 8 : 4;
    const void *sample = ezi ? (const char *)ezi + esz * (size_t)sample_index : nullptr;
    return fdtd::launch_fourier(dtype, nf, (size_t)nx * ny, cosv, sinv, ez, sample, ft, fdtd::as_stream(stream));
}

int fdtd2d_pmlparam(int dtype, int nx, int ny, int npml, const double *cubes, const fdtd_pmlayer *pml, void *stream) {
    FDTD_REQUIRE(nx >= 1 && ny >= 1 && npml >= 0 && 2 * npml <= (nx < ny ? nx : ny) && pml, "fdtd2d_pmlparam: npml=%d does not fit a %dx%d grid", npml, nx, ny);
    const void *vec[10] = {pml->fx1, pml->fx2, pml->fx3, pml->fy1, pml->fy2, pml->fy3, pml->gx2, pml->gx3, pml->gy2, pml->gy3};
    for (int k = 0; k < 10; ++k) FDTD_REQUIRE(vec[k] != nullptr, "fdtd2d_pmlparam: PML vector %d is null", k);
    const int n = nx > ny ? nx : ny;
    DISPATCH(dtype, (k_pmlparam<real><<<(n + 255) / 256, 256, 0, fdtd::as_stream(stream)>>>(nx, ny, npml, cubes,
        (real *)pml->fx1, (real *)pml->fx2, (real *)pml->fx3, (real *)pml->fy1, (real *)pml->fy2, (real *)pml->fy3,
        (real *)pml->gx2, (real *)pml->gx3, (real *)pml->gy2, (real *)pml->gy3)));
    FDTD_LAUNCH_CHECK("k_pmlparam");
    return FDTD_OK;
}

int fdtd2d_dielectric_cylinder(int dtype, int nx, int ny, int npml, int rgrid, double dt, double epsr, double sigma,
                               int row_lo, int row_hi, void *naz, void *nbz, void *stream) {
    FDTD_REQUIRE(nx >= 1 && ny >= 1 && npml >= 0 && row_lo >= 0 && row_hi <= nx && row_lo < row_hi && naz && nbz,
                 "fdtd2d_dielectric_cylinder: bad arguments");
    const dim3 grid((ny + BX - 1) / BX, (row_hi - row_lo + BY - 1) / BY);
    DISPATCH(dtype, (k_cylinder<real><<<grid, dim3(BX, BY), 0, fdtd::as_stream(stream)>>>(nx, ny, npml, rgrid, dt, epsr, sigma, row_lo, row_hi, (real *)naz, (real *)nbz)));
    FDTD_LAUNCH_CHECK("k_cylinder");
    return FDTD_OK;
}

int fdtd2d_hfield(int dtype, int nx, int ny, const fdtd_pmlayer *pml, const void *ez, void *ihx, void *ihy, void *hx,
                  void *hy, void *stream) {
    FDTD_REQUIRE(nx >= 2 && ny >= 2 && pml && ez && ihx && ihy && hx && hy, "fdtd2d_hfield: bad arguments");
    DISPATCH(dtype, (k_hfield<real><<<grid2d(nx, ny), dim3(BX, BY), 0, fdtd::as_stream(stream)>>>(nx, ny, Pml<real>(*pml), (const real *)ez, (real *)ihx, (real *)ihy, (real *)hx, (real *)hy)));
    FDTD_LAUNCH_CHECK("k_hfield");
    return FDTD_OK;
}

int fdtd2d_incthx(int dtype, int nx, int ny, int npml, const void *ezi, void *hx, void *stream) {
    FDTD_REQUIRE(tfsf_geometry_ok(nx, ny, npml) && ezi && hx, "fdtd2d_incthx: bad geometry nx=%d ny=%d npml=%d", nx, ny, npml);
    int n = nx - 2 * npml + 2;
    DISPATCH(dtype, (k_incthx<real><<<(n + 255) / 256, 256, 0, fdtd::as_stream(stream)>>>(nx, ny, npml, (const real *)ezi, (real *)hx)));
    FDTD_LAUNCH_CHECK("k_incthx");
    return FDTD_OK;
}

int fdtd2d_incthy(int dtype, int nx, int ny, int npml, const void *ezi, void *hy, void *stream) {
    FDTD_REQUIRE(tfsf_geometry_ok(nx, ny, npml) && ezi && hy, "fdtd2d_incthy: bad geometry nx=%d ny=%d npml=%d", nx, ny, npml);
    int n = ny - 2 * npml + 2;
    DISPATCH(dtype, (k_incthy<real><<<(n + 255) / 256, 256, 0, fdtd::as_stream(stream)>>>(nx, ny, npml, (const real *)ezi, (real *)hy)));
    FDTD_LAUNCH_CHECK("k_incthy");
    return FDTD_OK;
}

}  // extern "C"
