// Fused, temporally blocked 2D TM kernel ("march"): T full time steps (dfield+source+inctdz+efield+hfield+
// incthx+incthy, reference order fd2d/python/fd2d_3_4.py:268-277) per pass over HBM.
//
// Decomposition: one WARP owns a column strip of 32*V columns (V consecutive columns per lane, vector
// loads) and a chunk of rows, and marches down the rows.  The T time steps form a register pipeline:
// stage s holds the previous row (its new Dz/Ez and old H) and, when the next row arrives, finishes
//   D,E of the arriving row   (needs old Hy of the held row, Hx of the left column -> one warp shuffle)
//   H   of the held row       (needs Ez of the arriving row, Ez of the right column -> one warp shuffle)
// and hands the held row, now fully at time t+s+1, to stage s+1.  No shared memory, no block barrier:
// warps are independent; strips overlap by T columns and chunks by T rows on each side (recomputed).
// State is ping-ponged between two array sets, so a pass reads set A (time t) and writes set B (t+T);
// real HBM traffic is 48 B per cell per PASS instead of per step.
//
// Arithmetic is the reference's, operation by operation, no FMA contraction (-fmad=false) -> all six
// arrays are bit-identical to fd2d/program/fd2d_3_3.py run for the same number of steps.
#include "common.cuh"

namespace {

constexpr int TMAX = 8;          // deepest pipeline instantiated
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
    int tfsf, npml;
    const real *ezi_hist, *hxi_hist;           // [T][ny], [T][2]
    int src_i, src_j, src_hard;                // point source on dz (src_i < 0: none)
    double src[TMAX];
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

template <typename real, int V, bool LOSSY>
struct Row {           // one grid row as it travels between stages (state at one time level)
    real dz[V], hx[V], hy[V], ihx[V], ihy[V], naz[V], iz[V], nbz[V];
};

template <typename real, int V, bool LOSSY>
struct Held {          // the row a stage holds: D/E already advanced, H not yet
    real dz[V], ez[V], hx[V], hy[V], ihx[V], ihy[V], naz[V], iz[V], nbz[V];
};

template <typename real, int V, bool LOSSY>
__device__ __forceinline__ void load_row(const MarchParams<real> &p, int g, bool col_in, int jb,
                                         Row<real, V, LOSSY> &r) {
    if (g >= p.in_lo && g < p.in_hi && col_in) {
        const size_t off = (size_t)(g - p.row_base) * (size_t)p.ny + (size_t)jb;
        VecIO<real, V>::ld(p.in_dz + off, r.dz);
        VecIO<real, V>::ld(p.in_hx + off, r.hx);
        VecIO<real, V>::ld(p.in_hy + off, r.hy);
        VecIO<real, V>::ld(p.in_ihx + off, r.ihx);
        VecIO<real, V>::ld(p.in_ihy + off, r.ihy);
        VecIO<real, V>::ld(p.naz + off, r.naz);
        if (LOSSY) {
            VecIO<real, V>::ld(p.in_iz + off, r.iz);
            VecIO<real, V>::ld(p.nbz + off, r.nbz);
        }
    } else {
#pragma unroll
        for (int v = 0; v < V; ++v) {
            r.dz[v] = r.hx[v] = r.hy[v] = r.ihx[v] = r.ihy[v] = r.naz[v] = real(0);
            r.iz[v] = r.nbz[v] = real(0);
        }
    }
}

template <typename real, int V, int T, bool LOSSY>
__global__ void __launch_bounds__(MAX_WARPS * 32)
k_march(const __grid_constant__ MarchParams<real> p) {
    constexpr int W = 32 * V;            // columns per strip
    constexpr int USE = W - 2 * T;       // columns a strip produces
    constexpr unsigned FULL = 0xffffffffu;
    const real half = real(0.5);

    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= p.nstrips * p.nchunks) return;
    const int strip = w % p.nstrips;
    const int chunk = w / p.nstrips;

    const int c0 = strip * USE - T;                  // first column of the strip (halo included)
    const int jb = c0 + lane * V;                    // first column of this lane
    const bool col_in = (jb >= 0) && (jb + V <= p.ny);
    const bool col_store = col_in && (lane * V >= T) && (lane * V + V <= W - T);
    const int i0 = p.out_lo + chunk * p.chunk_rows;
    const int i1 = min(i0 + p.chunk_rows, p.out_hi);

    // per-column PML coefficients live in registers for the whole march
    real gy2[V], gy3[V], fy1[V], fy2[V], fy3[V];
    unsigned dmask = 0, hmask = 0;                   // bit v: D / H update applies to column jb+v
#pragma unroll
    for (int v = 0; v < V; ++v) {
        gy2[v] = gy3[v] = fy2[v] = fy3[v] = real(1);
        fy1[v] = real(0);
    }
    if (col_in) {
        VecIO<real, V>::ld(p.gy2 + jb, gy2);
        VecIO<real, V>::ld(p.gy3 + jb, gy3);
        VecIO<real, V>::ld(p.fy1 + jb, fy1);
        VecIO<real, V>::ld(p.fy2 + jb, fy2);
        VecIO<real, V>::ld(p.fy3 + jb, fy3);
#pragma unroll
        for (int v = 0; v < V; ++v) {
            if (jb + v >= 1) dmask |= 1u << v;
            if (jb + v <= p.ny - 2) hmask |= 1u << v;
        }
    }

    // warp-uniform "does this strip touch a special column" flags
    const int ja = p.npml - 1, jz = p.ny - p.npml;   // TFSF box edges along j: [ja, jz]
    const bool tf_cols = p.tfsf && ((ja - 1 >= c0 && ja - 1 < c0 + W) || (ja >= c0 && ja < c0 + W) ||
                                    (jz >= c0 && jz < c0 + W));
    const bool src_cols = (p.src_i >= 0) && (p.src_j >= c0 && p.src_j < c0 + W);
    const int ia = p.npml - 1, iz_ = p.nx - p.npml;  // TFSF box edges along i: [ia, iz_]

    Held<real, V, LOSSY> P[T];
#pragma unroll
    for (int s = 0; s < T; ++s)
#pragma unroll
        for (int v = 0; v < V; ++v) {
            P[s].dz[v] = P[s].ez[v] = P[s].hx[v] = P[s].hy[v] = P[s].ihx[v] = P[s].ihy[v] = real(0);
            P[s].naz[v] = P[s].iz[v] = P[s].nbz[v] = real(0);
        }

    Row<real, V, LOSSY> cur, nxt;
    load_row<real, V, LOSSY>(p, i0 - T, col_in, jb, nxt);

    for (int r = i0 - T; r < i1 + T; ++r) {
        cur = nxt;
        load_row<real, V, LOSSY>(p, r + 1 < i1 + T ? r + 1 : -1, col_in, jb, nxt);   // prefetch next row

        real ez_out[V];
#pragma unroll
        for (int s = 0; s < T; ++s) {
            const int rs = r - s;                    // global row carried by `cur` at this stage
            const int hr = rs - 1;                   // global row held by this stage
            const int rd = min(max(rs, 0), p.nx - 1);
            const int rh = min(max(hr, 0), p.nx - 1);
            const real gx2 = __ldg(p.gx2 + rd), gx3 = __ldg(p.gx3 + rd);
            const real fx1 = __ldg(p.fx1 + rh), fx2 = __ldg(p.fx2 + rh), fx3 = __ldg(p.fx3 + rh);
            const bool drow = (rs >= 1) && (rs < p.nx);
            const bool hrow = (hr >= 0) && (hr <= p.nx - 2);
            Held<real, V, LOSSY> &H = P[s];

            // ---- D of row rs:  dz = gx3*gy3*dz + gx2*gy2*0.5*(hy - hy[i-1] - hx + hx[j-1])
            real d[V], e[V], iznew[V];
            const real hx_left = __shfl_up_sync(FULL, cur.hx[V - 1], 1);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const real hxl = (v == 0) ? hx_left : cur.hx[v == 0 ? 0 : v - 1];
                const real curl = ((cur.hy[v] - H.hy[v]) - cur.hx[v]) + hxl;
                const real dn = ((gx3 * gy3[v]) * cur.dz[v]) + (((gx2 * gy2[v]) * half) * curl);
                d[v] = (drow && ((dmask >> v) & 1u)) ? dn : cur.dz[v];
            }
            if (src_cols && rs == p.src_i) {         // point source (after the stencil, before inctdz)
#pragma unroll
                for (int v = 0; v < V; ++v)
                    if (jb + v == p.src_j) d[v] = fdtd::inject<real>(d[v], p.src[s], p.src_hard);
            }
            if (tf_cols && rs >= ia && rs <= iz_) {  // inctdz: uses hxi of the previous step
                const real a = half * __ldg(p.hxi_hist + 2 * s), b = half * __ldg(p.hxi_hist + 2 * s + 1);
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    if (jb + v == ja) d[v] = d[v] + a;
                    if (jb + v == jz) d[v] = d[v] - b;
                }
            }
            // ---- E of row rs
#pragma unroll
            for (int v = 0; v < V; ++v) {
                if (LOSSY) {
                    e[v] = cur.naz[v] * (d[v] - cur.iz[v]);
                    iznew[v] = cur.iz[v] + cur.nbz[v] * e[v];
                } else {
                    e[v] = cur.naz[v] * d[v];
                    iznew[v] = real(0);
                }
            }
            // ---- H of the held row hr (needs ez[hr][j+1] and ez[rs][j])
            real hxn[V], hyn[V], ax[V], ay[V];
            const real ez_right = __shfl_down_sync(FULL, H.ez[0], 1);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const real er = (v == V - 1) ? ez_right : H.ez[v == V - 1 ? v : v + 1];
                const real cm = H.ez[v] - er;
                const real cn = H.ez[v] - e[v];
                const real sx = H.ihx[v] + cm;
                const real sy = H.ihy[v] + cn;
                const real hx2 = (fy3[v] * H.hx[v]) + (fy2[v] * ((half * cm) + (fx1 * sx)));
                const real hy2 = (fx3 * H.hy[v]) - (fx2 * ((half * cn) + (fy1[v] * sy)));
                const bool up = hrow && ((hmask >> v) & 1u);
                ax[v] = up ? sx : H.ihx[v];
                ay[v] = up ? sy : H.ihy[v];
                hxn[v] = up ? hx2 : H.hx[v];
                hyn[v] = up ? hy2 : H.hy[v];
            }
            if (p.tfsf) {
                if (tf_cols && hr >= ia && hr <= iz_) {      // incthx
                    const real *ez_i = p.ezi_hist + (size_t)s * p.ny;
                    const real a = half * __ldg(ez_i + ja), b = half * __ldg(ez_i + jz);
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        if (jb + v == ja - 1) hxn[v] = hxn[v] + a;
                        if (jb + v == jz) hxn[v] = hxn[v] - b;
                    }
                }
                if (hr == ia - 1 || hr == iz_) {             // incthy (two rows of the whole grid)
                    const real *ez_i = p.ezi_hist + (size_t)s * p.ny;
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        const int j = jb + v;
                        if (j >= ja && j <= jz) {
                            const real h = half * __ldg(ez_i + j);
                            if (hr == ia - 1) hyn[v] = hyn[v] - h;
                            if (hr == iz_) hyn[v] = hyn[v] + h;
                        }
                    }
                }
            }
            // ---- hand the finished row to the next stage, keep the arriving one
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const real odz = H.dz[v], onaz = H.naz[v], oiz = H.iz[v], onbz = H.nbz[v];
                ez_out[v] = H.ez[v];
                H.dz[v] = d[v];   H.ez[v] = e[v];
                H.hx[v] = cur.hx[v];  H.hy[v] = cur.hy[v];
                H.ihx[v] = cur.ihx[v];  H.ihy[v] = cur.ihy[v];
                H.naz[v] = cur.naz[v];
                if (LOSSY) { H.iz[v] = iznew[v]; H.nbz[v] = cur.nbz[v]; }
                cur.dz[v] = odz;  cur.hx[v] = hxn[v];  cur.hy[v] = hyn[v];
                cur.ihx[v] = ax[v];  cur.ihy[v] = ay[v];  cur.naz[v] = onaz;
                if (LOSSY) { cur.iz[v] = oiz; cur.nbz[v] = onbz; }
            }
        }

        // `cur` is now row r-T at time t+T
        const int ro = r - T;
        if (ro >= i0 && ro < i1 && col_store) {
            const size_t off = (size_t)(ro - p.row_base) * (size_t)p.ny + (size_t)jb;
            VecIO<real, V>::st(p.out_dz + off, cur.dz);
            VecIO<real, V>::st(p.out_ez + off, ez_out);
            VecIO<real, V>::st(p.out_hx + off, cur.hx);
            VecIO<real, V>::st(p.out_hy + off, cur.hy);
            VecIO<real, V>::st(p.out_ihx + off, cur.ihx);
            VecIO<real, V>::st(p.out_ihy + off, cur.ihy);
            if (LOSSY) VecIO<real, V>::st(p.out_iz + off, cur.iz);
        }
    }
}

// ---- incident line: T steps of the 1D auxiliary FDTD (ezinct ... hxinct), recording what the 2D pass
// needs: ezi after ezinct+source of every sub-step, and hxi[npml-2], hxi[ny-npml] BEFORE hxinct.
struct SrcTable { double v[TMAX]; };

template <typename real>
__global__ void k_incident_line(int ny, int npml, int T, real *ezi, real *hxi, real *bc, real *ezi_hist,
                                real *hxi_hist, const SrcTable src) {
    for (int s = 0; s < T; ++s) {
        for (int j = 1 + threadIdx.x; j < ny; j += blockDim.x) ezi[j] = ezi[j] + real(0.5) * (hxi[j - 1] - hxi[j]);
        __syncthreads();
        if (threadIdx.x == 0) {
            real e1 = ezi[1], b0 = bc[0], b1 = bc[1];
            ezi[0] = b0; bc[0] = b1; bc[1] = e1;
            real e2 = ezi[ny - 2], b3 = bc[3], b2 = bc[2];
            ezi[ny - 1] = b3; bc[3] = b2; bc[2] = e2;
            ezi[3] = static_cast<real>(src.v[s]);
            hxi_hist[2 * s] = hxi[npml - 2];
            hxi_hist[2 * s + 1] = hxi[ny - npml];
        }
        __syncthreads();
        for (int j = threadIdx.x; j < ny; j += blockDim.x) {
            const real e = ezi[j];
            ezi_hist[(size_t)s * ny + j] = e;
            if (j < ny - 1) hxi[j] = hxi[j] + real(0.5) * (e - ezi[j + 1]);
        }
        __syncthreads();
    }
}

int g_force_v = 0;          // test / tuning hooks (fdtd2d_tune)
int g_chunk_rows = 0;
int g_warps = 0;

template <typename real, int V, int T>
int launch_march(const MarchParams<real> &mp, bool lossy, cudaStream_t st) {
    const int nw = mp.nstrips * mp.nchunks;
    const int warps = (g_warps >= 1 && g_warps <= MAX_WARPS) ? g_warps : 4;
    const int grid = (nw + warps - 1) / warps;
    if (lossy) k_march<real, V, T, true><<<grid, warps * 32, 0, st>>>(mp);
    else       k_march<real, V, T, false><<<grid, warps * 32, 0, st>>>(mp);
    FDTD_LAUNCH_CHECK("k_march");
    return FDTD_OK;
}

template <typename real, int V>
int launch_march_T(int T, const MarchParams<real> &mp, bool lossy, cudaStream_t st) {
    switch (T) {
        case 1: return launch_march<real, V, 1>(mp, lossy, st);
        case 2: return launch_march<real, V, 2>(mp, lossy, st);
        case 3: return launch_march<real, V, 3>(mp, lossy, st);
        case 4: return launch_march<real, V, 4>(mp, lossy, st);
        default: fdtd::set_error("unsupported time-block depth %d for vector width %d", T, V); return FDTD_EUNSUPPORTED;
    }
}

// vector width: widest V dividing ny and T (strip origin c0 = strip*(32V-2T) - T must be V-aligned)
template <typename real> int pick_v(int ny, int T);
template <> int pick_v<float>(int ny, int T) {
    if (ny % 4 == 0 && T % 4 == 0) return 4;
    if (ny % 2 == 0 && T % 2 == 0) return 2;
    return 1;
}
template <> int pick_v<double>(int ny, int T) { return (ny % 2 == 0 && T % 2 == 0) ? 2 : 1; }


template <typename real>
int advance(const fdtd2d_problem *q, int cur, int nsteps, const double *src, int tblock, cudaStream_t st,
            int *cur_out) {
    const bool lossy = (q->flags & FDTD_LOSSY) != 0, tfsf = (q->flags & FDTD_TFSF) != 0;
    int done = 0;
    while (done < nsteps) {
        const int T = min(tblock, nsteps - done);
        const int rem = nsteps - done - T;              // steps still to come after this pass
        MarchParams<real> mp;
        void *const *in = q->state[cur];
        void *const *out = q->state[cur ^ 1];
        mp.in_dz = (const real *)in[FDTD2D_DZ];   mp.in_hx = (const real *)in[FDTD2D_HX];
        mp.in_hy = (const real *)in[FDTD2D_HY];   mp.in_ihx = (const real *)in[FDTD2D_IHX];
        mp.in_ihy = (const real *)in[FDTD2D_IHY]; mp.in_iz = (const real *)in[FDTD2D_IZ];
        mp.out_dz = (real *)out[FDTD2D_DZ];   mp.out_ez = (real *)out[FDTD2D_EZ];
        mp.out_hx = (real *)out[FDTD2D_HX];   mp.out_hy = (real *)out[FDTD2D_HY];
        mp.out_ihx = (real *)out[FDTD2D_IHX]; mp.out_ihy = (real *)out[FDTD2D_IHY];
        mp.out_iz = (real *)out[FDTD2D_IZ];
        mp.naz = (const real *)q->md.naz;  mp.nbz = (const real *)q->md.nbz;
        mp.gx2 = (const real *)q->pml.gx2; mp.gx3 = (const real *)q->pml.gx3;
        mp.fx1 = (const real *)q->pml.fx1; mp.fx2 = (const real *)q->pml.fx2; mp.fx3 = (const real *)q->pml.fx3;
        mp.gy2 = (const real *)q->pml.gy2; mp.gy3 = (const real *)q->pml.gy3;
        mp.fy1 = (const real *)q->pml.fy1; mp.fy2 = (const real *)q->pml.fy2; mp.fy3 = (const real *)q->pml.fy3;
        mp.nx = q->nx; mp.ny = q->ny; mp.row_base = q->row_base;
        mp.in_lo = max(q->row_base, 0);
        mp.in_hi = min(q->row_base + q->rows_alloc, q->nx);
        mp.out_lo = max(q->row_lo - rem, mp.in_lo);
        mp.out_hi = min(q->row_hi + rem, mp.in_hi);
        mp.tfsf = tfsf; mp.npml = q->npml;
        mp.ezi_hist = (const real *)q->ezi_hist; mp.hxi_hist = (const real *)q->hxi_hist;
        mp.src_i = tfsf ? -1 : q->src_i; mp.src_j = q->src_j; mp.src_hard = q->src_hard;
        for (int s = 0; s < TMAX; ++s) mp.src[s] = (src && s < T) ? src[done + s] : 0.0;

        int V = g_force_v ? g_force_v : pick_v<real>(q->ny, T);
        if (q->ny % V != 0 || T % V != 0) V = 1;
        const int use = 32 * V - 2 * T;
        mp.nstrips = (q->ny + use - 1) / use;
        const int rows = mp.out_hi - mp.out_lo;
        int chunk = g_chunk_rows;
        if (chunk <= 0) {
            // enough warps for ~3 waves at 16 warps/SM, but keep the 2T-row recompute overhead <= ~6%
            const long want = 3L * fdtd::sm_count() * 16;
            long nchunks = (want + mp.nstrips - 1) / mp.nstrips;
            chunk = (int)((rows + nchunks - 1) / max(nchunks, 1L));
            chunk = max(chunk, 32 * T);
        }
        chunk = max(1, min(chunk, rows));
        mp.chunk_rows = chunk;
        mp.nchunks = (rows + chunk - 1) / chunk;

        if (tfsf) {
            SrcTable tab;
            for (int s = 0; s < TMAX; ++s) tab.v[s] = mp.src[s];
            k_incident_line<real><<<1, 1024, 0, st>>>(q->ny, q->npml, T, (real *)q->ezi, (real *)q->hxi, (real *)q->bc,
                                                      (real *)q->ezi_hist, (real *)q->hxi_hist, tab);
            FDTD_LAUNCH_CHECK("k_incident_line");
        }
        int rc;
        if constexpr (sizeof(real) == 4) {
            if (V == 4) rc = launch_march_T<real, 4>(T, mp, lossy, st);
            else if (V == 2) rc = launch_march_T<real, 2>(T, mp, lossy, st);
            else rc = launch_march_T<real, 1>(T, mp, lossy, st);
        } else {
            if (V == 2) rc = launch_march_T<real, 2>(T, mp, lossy, st);
            else rc = launch_march_T<real, 1>(T, mp, lossy, st);
        }
        if (rc != FDTD_OK) return rc;
        cur ^= 1;
        done += T;
    }
    *cur_out = cur;
    return FDTD_OK;
}

}  // namespace

extern "C" {

int fdtd2d_max_tblock(int dtype, int ny) {
    (void)ny;
    return (dtype == FDTD_F32 || dtype == FDTD_F64) ? 4 : 0;
}

// tuning / test hook (not part of the reference-facing surface): force the vector width and rows per chunk
int fdtd2d_tune(int force_v, int chunk_rows, int warps_per_cta) {
    g_force_v = force_v;
    g_chunk_rows = chunk_rows;
    g_warps = warps_per_cta;
    return FDTD_OK;
}

int fdtd2d_advance(const fdtd2d_problem *q, int cur, int nsteps, const double *src, int tblock, void *stream,
                   int *cur_out) {
    FDTD_REQUIRE(q && cur_out, "fdtd2d_advance: null problem / cur_out");
    FDTD_REQUIRE(cur == 0 || cur == 1, "fdtd2d_advance: cur must be 0 or 1");
    FDTD_REQUIRE(q->nx >= 2 && q->ny >= 2, "fdtd2d_advance: grid %dx%d too small", q->nx, q->ny);
    FDTD_REQUIRE(tblock >= 1 && tblock <= 4, "fdtd2d_advance: tblock %d outside [1, 4]", tblock);
    FDTD_REQUIRE(nsteps >= 0, "fdtd2d_advance: nsteps < 0");
    FDTD_REQUIRE(q->row_lo >= 0 && q->row_hi <= q->nx && q->row_lo < q->row_hi, "fdtd2d_advance: bad owned rows [%d,%d)", q->row_lo, q->row_hi);
    FDTD_REQUIRE(q->row_base <= q->row_lo && q->row_base + q->rows_alloc >= q->row_hi, "fdtd2d_advance: owned rows outside the stored rows");
    {   // every row within nsteps of the owned range (clipped to the grid) must be stored
        const int need_lo = q->row_lo - nsteps > 0 ? q->row_lo - nsteps : 0;
        const int need_hi = q->row_hi + nsteps < q->nx ? q->row_hi + nsteps : q->nx;
        FDTD_REQUIRE(q->row_base <= need_lo && q->row_base + q->rows_alloc >= need_hi,
                     "fdtd2d_advance: %d steps need ghost rows [%d,%d) but only [%d,%d) are stored", nsteps, need_lo,
                     need_hi, q->row_base, q->row_base + q->rows_alloc);
    }
    const bool lossy = (q->flags & FDTD_LOSSY) != 0, tfsf = (q->flags & FDTD_TFSF) != 0;
    for (int s = 0; s < 2; ++s)
        for (int f = 0; f < FDTD2D_NFIELDS; ++f) {
            if (f == FDTD2D_IZ && !lossy) continue;
            FDTD_REQUIRE(q->state[s][f] && fdtd::aligned16(q->state[s][f]), "fdtd2d_advance: state[%d][%d] null or not 16-byte aligned", s, f);
        }
    FDTD_REQUIRE(q->md.naz && fdtd::aligned16(q->md.naz), "fdtd2d_advance: naz null or misaligned");
    FDTD_REQUIRE(!lossy || (q->md.nbz && fdtd::aligned16(q->md.nbz)), "fdtd2d_advance: FDTD_LOSSY needs nbz");
    const void *vec[10] = {q->pml.fx1, q->pml.fx2, q->pml.fx3, q->pml.fy1, q->pml.fy2,
                           q->pml.fy3, q->pml.gx2, q->pml.gx3, q->pml.gy2, q->pml.gy3};
    for (int k = 0; k < 10; ++k) FDTD_REQUIRE(vec[k] && fdtd::aligned16(vec[k]), "fdtd2d_advance: PML vector %d null or misaligned", k);
    if (tfsf) {
        FDTD_REQUIRE(q->npml >= 2 && 2 * q->npml <= q->nx && 2 * q->npml <= q->ny && q->ny >= 8, "fdtd2d_advance: TFSF needs 2 <= npml <= min(nx,ny)/2");
        FDTD_REQUIRE(q->ezi && q->hxi && q->bc && q->ezi_hist && q->hxi_hist, "fdtd2d_advance: TFSF buffers missing");
        FDTD_REQUIRE(src != nullptr || nsteps == 0, "fdtd2d_advance: TFSF needs a source table");
    } else if (q->src_i >= 0) {
        FDTD_REQUIRE(q->src_i < q->nx && q->src_j >= 0 && q->src_j < q->ny && src, "fdtd2d_advance: bad point source");
    }
    if (q->dtype == FDTD_F32) return advance<float>(q, cur, nsteps, src, tblock, fdtd::as_stream(stream), cur_out);
    if (q->dtype == FDTD_F64) return advance<double>(q, cur, nsteps, src, tblock, fdtd::as_stream(stream), cur_out);
    fdtd::set_error("fdtd2d_advance: unknown dtype %d", q->dtype);
    return FDTD_EINVAL;
}

}  // extern "C"
