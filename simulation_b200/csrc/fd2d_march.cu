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
#include "fd2d_march.cuh"

#include <mutex>

namespace {

using namespace fdtd_march;


// The march of one warp over its (strip, chunk).  Rows are staged through a per-lane ring of RING rows in
// shared memory filled by cp.async (each lane reads back only the 16 B it copied itself: no barrier, no
// register cost), keeping RING-1 rows x 6 arrays in flight per warp to cover the HBM latency.
// (RING itself is declared in fd2d_march.cuh: the host's row classification needs the fetch run-ahead too.)

// Compile-time shape of one instantiation: register row sets, ring slots, shared memory per warp.
// NAZR (deep interior pipelines, T >= 7): a row set of 7 arrays x 4 columns x 9 sets does not fit the register
// file, so naz -- read once per stage, never written -- stays in shared memory: its own ring of 2*(T+1) rows,
// addressed with compile-time slots (the row loop is unrolled T+1 times; the two halves alternate per trip).
template <typename real, int V, int T, int MODE, bool FAST>
struct MarchShape {
    static constexpr bool LOSSY = (MODE & 1) != 0, DFT = (MODE & 2) != 0;
    static constexpr bool NAZR = FAST && !LOSSY && !DFT && T >= 7;
    static constexpr bool PACKED = FAST && !DFT && sizeof(real) == 4 && V % 2 == 0;   // FADD2 / FFMA2 stage
    static constexpr int NS = T + 1;                                   // register row sets
    static constexpr int NARR = (LOSSY ? 8 : 6) + (DFT ? 2 * NFMAX : 0) - (NAZR ? 1 : 0);   // arrays per main-ring row
    static constexpr int LB = V * (int)sizeof(real);                   // bytes per lane per array row
    static constexpr int ROWB = 32 * LB;                               // bytes per array row (per warp)
    static constexpr int SLOT = NARR * ROWB;                           // main-ring bytes per row
    static constexpr int NAZ_ROWS = NAZR ? 2 * NS : 0;
    static constexpr int WARP_SMEM = RING * SLOT + NAZ_ROWS * ROWB;
};

template <typename real, int V, int T, int MODE, bool FAST>
__device__ __forceinline__ void march_body(const MarchParams<real> &p, const int strip, const int i0, const int i1,
                                           const int lane, unsigned char *const ring) {
    using Shape = MarchShape<real, V, T, MODE, FAST>;
    constexpr bool LOSSY = Shape::LOSSY, DFT = Shape::DFT, NAZR = Shape::NAZR;
    constexpr int W = 32 * V;            // columns per strip
    constexpr int HALO = ((T + V - 1) / V) * V;   // recomputed columns per side: >= T, multiple of V (aligned vectors)
    constexpr int USE = W - 2 * HALO;    // columns a strip produces
    constexpr int NS = Shape::NS;
    constexpr int LB = Shape::LB, ROWB = Shape::ROWB, SLOT = Shape::SLOT;

    const int c0 = strip * USE - HALO;               // first column of the strip (halo included)
    const int jb = c0 + lane * V;                    // first column of this lane
    const bool col_in = FAST || ((jb >= 0) && (jb + V <= p.ny));
    const bool col_store = col_in && (lane * V >= HALO) && (lane * V + V <= W - HALO);

    ColCoef<real, V> c;
    c.dmask = c.hmask = 0;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        c.gy2[v] = c.gy3[v] = c.fy2[v] = c.fy3[v] = real(1);
        c.fy1[v] = real(0);
    }
    if (!FAST && col_in) {
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
    // warp-uniform "does this strip touch a special column" flags (careful path only)
    const int ja = p.npml - 1, jz = p.ny - p.npml;
    const bool tf_cols = !FAST && p.tfsf && ((ja - 1 >= c0 && ja - 1 < c0 + W) || (ja >= c0 && ja < c0 + W) ||
                                             (jz >= c0 && jz < c0 + W));
    const bool src_cols = !FAST && (p.src_i >= 0) && (p.src_j >= c0 && p.src_j < c0 + W);

    RowSet<real, V> S[NS];
#pragma unroll
    for (int k = 0; k < NS; ++k)
#pragma unroll
        for (int v = 0; v < V; ++v) {
            S[k].dz[v] = S[k].ez[v] = S[k].hx[v] = S[k].hy[v] = S[k].ihx[v] = S[k].ihy[v] = real(0);
            S[k].naz[v] = S[k].iz[v] = S[k].nbz[v] = real(0);
#pragma unroll
            for (int f = 0; f < NFMAX; ++f) S[k].racc[f][v] = S[k].iacc[f][v] = real(0);
        }

    unsigned char *const lane_ring = ring + lane * LB;
    unsigned char *const lane_naz = ring + RING * SLOT + lane * LB;      // NAZR: this lane's column of the naz ring
    const int r_begin = i0 - T, r_end = i1 + T;      // rows fed to stage 0: [r_begin, r_end)
    // element offsets (row-major, local array rows) of the row fetched next / stored next; advanced by ny per row
    long long off_f = (long long)(r_begin - p.row_base) * p.ny + jb;
    long long off_s = (long long)(r_begin - T - p.row_base) * p.ny + jb;
    int g_f = r_begin;                                // global row fetched next
    // asynchronous copies of global row g_f into ring slot k (zero fill outside the stored rows / the grid;
    // FAST warps only ever touch stored interior rows, so their copies are unconditional)
    auto fetch = [&](const int k, const int naz_off) {
        const bool ok = FAST || ((g_f >= p.in_lo) && (g_f < p.in_hi) && (g_f < r_end) && col_in);
        const long long off = ok ? off_f : 0;
        const int nb = ok ? LB : 0;
        unsigned char *dst = lane_ring + k * SLOT;
        cp_async<LB>(dst + 0 * 32 * LB, p.in_dz + off, nb);
        cp_async<LB>(dst + 1 * 32 * LB, p.in_hx + off, nb);
        cp_async<LB>(dst + 2 * 32 * LB, p.in_hy + off, nb);
        cp_async<LB>(dst + 3 * 32 * LB, p.in_ihx + off, nb);
        cp_async<LB>(dst + 4 * 32 * LB, p.in_ihy + off, nb);
        if (NAZR) cp_async<LB>(lane_naz + naz_off, p.naz + off, nb);
        else cp_async<LB>(dst + 5 * 32 * LB, p.naz + off, nb);
        if (LOSSY) {
            cp_async<LB>(dst + 6 * 32 * LB, p.in_iz + off, nb);
            cp_async<LB>(dst + 7 * 32 * LB, p.nbz + off, nb);
        }
        if (DFT) {   // accumulators are read (and later written) by the OWNER of a cell only: no halo reads, no race
            const bool own = (g_f >= i0) && (g_f < i1) && col_store;
            const long long offa = own ? off_f : 0;
            const int na = own ? LB : 0;
            constexpr int A0 = LOSSY ? 8 : 6;
#pragma unroll
            for (int f = 0; f < NFMAX; ++f) {
                const bool on = f < p.nf;
                cp_async<LB>(dst + (A0 + 2 * f) * 32 * LB, p.r_pt + (on ? f * p.dft_plane + offa : 0), on ? na : 0);
                cp_async<LB>(dst + (A0 + 2 * f + 1) * 32 * LB, p.i_pt + (on ? f * p.dft_plane + offa : 0), on ? na : 0);
            }
        }
        cp_async_commit();
        off_f += p.ny;
        ++g_f;
    };
    auto take = [&](const int k, RowSet<real, V> &row) {
        const unsigned char *src = lane_ring + k * SLOT;
        lds_vec<real, V>(src + 0 * 32 * LB, row.dz);
        lds_vec<real, V>(src + 1 * 32 * LB, row.hx);
        lds_vec<real, V>(src + 2 * 32 * LB, row.hy);
        lds_vec<real, V>(src + 3 * 32 * LB, row.ihx);
        lds_vec<real, V>(src + 4 * 32 * LB, row.ihy);
        if (!NAZR) lds_vec<real, V>(src + 5 * 32 * LB, row.naz);
        if (LOSSY) {
            lds_vec<real, V>(src + 6 * 32 * LB, row.iz);
            lds_vec<real, V>(src + 7 * 32 * LB, row.nbz);
        }
        if (DFT) {
            constexpr int A0 = LOSSY ? 8 : 6;
#pragma unroll
            for (int f = 0; f < NFMAX; ++f) {
                lds_vec<real, V>(src + (A0 + 2 * f) * 32 * LB, row.racc[f]);
                lds_vec<real, V>(src + (A0 + 2 * f + 1) * 32 * LB, row.iacc[f]);
            }
        }
    };

    if (NAZR) {        // rows above the chunk (pipeline warm-up) read naz slots no copy has filled yet
#pragma unroll
        for (int k = 0; k < Shape::NAZ_ROWS; ++k)
#pragma unroll
            for (int b = 0; b < LB; b += 4) *reinterpret_cast<float *>(lane_naz + k * ROWB + b) = 0.f;
    }
#pragma unroll
    for (int k = 0; k < RING - 1; ++k) fetch(k, k * ROWB);
    int slot = 0;                                     // ring slot of the row consumed next

    // packed kernels keep values in register PAIRS: storing a row as two 64-bit halves spares the moves that
    // assembling aligned quads for 128-bit stores would cost (same sectors, same bytes)
    auto st_vec = [&](real *dst, const real (&d)[V]) {
        if constexpr (Shape::PACKED && V == 4) {
            *reinterpret_cast<float2 *>(dst) = make_float2(d[0], d[1]);
            *reinterpret_cast<float2 *>(dst + 2) = make_float2(d[2], d[3]);
        } else {
            VecIO<real, V>::st(dst, d);
        }
    };
    auto store_row = [&](const RowSet<real, V> &O, const int ro, const int naz_off) {
        if (ro >= i0 && ro < i1 && col_store) {
            st_vec(p.out_dz + off_s, O.dz);
            if (p.write_ez) {
                if constexpr (NAZR) {     // Ez is not carried by the sets: the last pass evaluates naz*dz once more
                    real nz[V], e[V];
                    lds_vec<real, V>(lane_naz + naz_off, nz);
#pragma unroll
                    for (int v = 0; v < V; ++v) e[v] = nz[v] * O.dz[v];
                    st_vec(p.out_ez + off_s, e);
                } else {
                    st_vec(p.out_ez + off_s, O.ez);
                }
            }
            st_vec(p.out_hx + off_s, O.hx);
            st_vec(p.out_hy + off_s, O.hy);
            st_vec(p.out_ihx + off_s, O.ihx);
            st_vec(p.out_ihy + off_s, O.ihy);
            if (LOSSY) st_vec(p.out_iz + off_s, O.iz);
            if (DFT) {
#pragma unroll
                for (int f = 0; f < NFMAX; ++f)
                    if (f < p.nf) {
                        VecIO<real, V>::st(p.r_pt + f * p.dft_plane + off_s, O.racc[f]);
                        VecIO<real, V>::st(p.i_pt + f * p.dft_plane + off_s, O.iacc[f]);
                    }
            }
            // halo exchange fused into the pass: peer stores over NVLink, row by row.  Only the careful kernel
            // pushes: the host lists every chunk that owns a pushed row (or reads a ghost row) as special.
            if (!FAST && p.push) push_row<real, V, LOSSY>(p, off_s, ro, O.dz, O.hx, O.hy, O.ihx, O.ihy, O.iz);
        }
        off_s += p.ny;
    };

    if constexpr (FAST) {
        // interior: row loop unrolled NS times, the register sets rotate through the roles (no moves)
        float2 negzero;
        {
            const unsigned long long z = p.negzero2;
            negzero = make_float2(__uint_as_float((unsigned)z), __uint_as_float((unsigned)(z >> 32)));
        }
        int nz_here = 0, nz_other = NS * ROWB;        // NAZR: byte offsets of this trip's / the other half of the naz ring
#pragma unroll 1                                      // the body is T*(T+1) stages already: a second copy only costs i-cache
        for (int r = r_begin; r < r_end; r += NS) {
#pragma unroll
            for (int u = 0; u < NS; ++u) {
                const int rr = r + u;                 // global row arriving at stage 0 (may overrun r_end)
                cp_async_wait<RING - 2>();            // the oldest of the RING-1 pending rows has landed
                take(slot, S[u]);
                // refill the slot consumed one sub-iteration ago with row rr + RING - 1
                fetch(slot == 0 ? RING - 1 : slot - 1,
                      (u + RING - 1 < NS) ? nz_here + (u + RING - 1) * ROWB : nz_other + (u + RING - 1 - NS) * ROWB);
                slot = (slot + 1 == RING) ? 0 : slot + 1;
#pragma unroll
                for (int s = 0; s < T; ++s) {     // row rr-s: fetched this trip (u >= s) or by the previous one
                    const void *nzA = lane_naz + ((u >= s) ? nz_here + (u - s) * ROWB : nz_other + (NS + u - s) * ROWB);
                    const void *nzH = lane_naz + ((u >= s + 1) ? nz_here + (u - s - 1) * ROWB : nz_other + (NS + u - s - 1) * ROWB);
                    if constexpr (Shape::PACKED)
                        march_stage_pk<V, NAZR, LOSSY>(S[(u - s + 2 * NS) % NS], S[(u - s - 1 + 2 * NS) % NS], negzero, nzA, nzH);
                    else
                        march_stage<real, V, MODE, FAST, NAZR>(p, c, S[(u - s + 2 * NS) % NS], S[(u - s - 1 + 2 * NS) % NS], rr - s,
                                                                s, jb, tf_cols, src_cols, nzA, nzH);
                }
                // the set held by the last stage: row rr-T at time t+T
                store_row(S[(u + 1) % NS], rr - T, (u >= T) ? nz_here + (u - T) * ROWB : nz_other + (NS + u - T) * ROWB);
            }
            const int t = nz_here; nz_here = nz_other; nz_other = t;
        }
    } else {
        // edges: compact code matters more than instruction count (this kernel is a fraction of a wave and
        // shares the instruction cache with the interior kernel): one rolled row loop, the sets are shifted
        // by register copies -- S[s] arrives at stage s, S[s+1] is held by it, S[T] leaves.
#pragma unroll 1
        for (int rr = r_begin; rr < r_end; ++rr) {
            cp_async_wait<RING - 2>();
            take(slot, S[0]);
            fetch(slot == 0 ? RING - 1 : slot - 1, 0);
            slot = (slot + 1 == RING) ? 0 : slot + 1;
#pragma unroll
            for (int s = 0; s < T; ++s)
                march_stage<real, V, MODE, FAST>(p, c, S[s], S[s + 1], rr - s, s, jb, tf_cols, src_cols);
            store_row(S[T], rr - T, 0);
#pragma unroll
            for (int k = T; k >= 1; --k) S[k] = S[k - 1];
        }
    }
    cp_async_wait<0>();
}


// Two launches per pass share this kernel template:
//   FAST    : interior warps -- (ordinary strips) x (ordinary chunks): every column (halo included) is an
//             ordinary cell and every row touched (warm-up, drain and unroll overrun included) is an ordinary
//             stored row, so the body carries no masks, clamps or TFSF / source cells;
//   careful : the listed special strips x all chunks, plus ordinary strips x the listed special chunks
//             (grid edges, PEC row / column, TFSF box edges, the point source); or everything (all_careful).
template <typename real, int V, int T, int MODE, bool FAST>
__global__ void __launch_bounds__(MAX_WARPS * 32)
k_march(const __grid_constant__ MarchParams<real> p, const int all_careful) {
    extern __shared__ __align__(16) unsigned char ring_smem[];
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    unsigned char *const ring = ring_smem + (size_t)(threadIdx.x >> 5) * MarchShape<real, V, T, MODE, FAST>::WARP_SMEM;
    int strip, i0, i1;
    if (!decode_item<FAST>(p, w, all_careful, V, T, (MODE & 1) != 0, strip, i0, i1)) return;
    if (!FAST) halo_wait(p, lane);
    march_body<real, V, T, MODE, FAST>(p, strip, i0, i1, lane, ring);
    if (!FAST) halo_signal(p, lane);
}

// ---- incident line: T steps of the 1D auxiliary FDTD (ezinct ... hxinct), recording what the 2D pass
// needs: ezi after ezinct+source of every sub-step, and hxi[npml-2], hxi[ny-npml] BEFORE hxinct.
struct SrcTable {
    double v[TMAX];
    int nf;                                   // running DFT of the source sample ezi[6] (0 = off)
    double c[TMAX][NFMAX], s[TMAX][NFMAX];
};

template <typename real>
__global__ void k_incident_line(int ny, int npml, int T, real *ezi, real *hxi, real *bc, real *ezi_hist,
                                real *hxi_hist, real *r_in, real *i_in, const SrcTable src) {
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
            for (int f = 0; f < src.nf; ++f) {        // fourier: the source sample, after ezinct + source
                const double e = static_cast<double>(ezi[6]);
                r_in[f] = static_cast<real>(static_cast<double>(r_in[f]) + src.c[s][f] * e);
                i_in[f] = static_cast<real>(static_cast<double>(i_in[f]) - src.s[s][f] * e);
            }
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

// counts entries of the promised identity ranges that are not exactly the identity coefficient set
template <typename real>
__global__ void k_check_identity(const real *f1, const real *f2, const real *f3, const real *g2, const real *g3, int lo,
                                 int hi, unsigned long long *bad) {
    const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    const bool ok = f1[i] == real(0) && !signbit(f1[i]) && f2[i] == real(1) && f3[i] == real(1) && g2[i] == real(1) &&
                    g3[i] == real(1);
    if (!ok) atomicAdd(bad, 1ull);
}

// counts stored cells OUTSIDE the promised box whose nbz is not 0 or whose iz (either state set) is not +0
template <typename real>
__global__ void k_check_lossless_outside(const real *nbz, const real *iz0, const real *iz1, int ny, int row_base, int rows,
                                         int rlo, int rhi, int clo, int chi, unsigned long long *bad) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ny) return;
    const real zero = real(0);
    for (int r = blockIdx.y; r < rows; r += gridDim.y) {
        const int i = row_base + r;
        if (i >= rlo && i < rhi && j >= clo && j < chi) continue;
        const size_t n = (size_t)r * ny + j;
        const bool ok = nbz[n] == zero && iz0[n] == zero && !signbit(iz0[n]) && iz1[n] == zero && !signbit(iz1[n]);
        if (!ok) atomicAdd(bad, 1ull);
    }
}

}  // namespace

namespace fdtd_march {

Tuning g_tune;               // test / tuning hooks (fdtd2d_tune, fdtd2d_tune2)

// one high-priority side stream + fork/join events per (device, launch stream), created on first use: callers that
// drive several launch streams at once (Fdtd2D.run_streamed) must not queue behind each other's edge kernels.
// The table is shared by every host thread: guarded by a mutex (entries are never removed, so the returned pointer
// stays valid).
SideStream *side_stream(cudaStream_t launch) {
    struct Entry { int dev; cudaStream_t launch; SideStream side; };
    constexpr int CAP = 64;
    static Entry table[CAP];
    static int used = 0;
    static std::mutex guard;
    std::lock_guard<std::mutex> lock(guard);
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    for (int k = 0; k < used; ++k)
        if (table[k].dev == dev && table[k].launch == launch) return &table[k].side;
    if (used == CAP) return nullptr;                    // callers fall back to the in-order launch
    Entry &e = table[used];
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithPriority(&e.side.stream, cudaStreamNonBlocking, hi) != cudaSuccess) return nullptr;
    if (cudaStreamCreateWithPriority(&e.side.backfill, cudaStreamNonBlocking, lo) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&e.side.fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&e.side.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    e.dev = dev;
    e.launch = launch;
    ++used;
    return &e.side;
}

}  // namespace fdtd_march

namespace {

template <typename real, int V, int T, int MODE, bool FAST>
int launch_one(const MarchParams<real> &mp, int items, int all_careful, cudaStream_t st) {
    if (items <= 0) return FDTD_OK;
    const size_t per_warp = MarchShape<real, V, T, MODE, FAST>::WARP_SMEM;
    // 8 independent warps per CTA on big grids; fewer when there are not enough warps to fill every SM
    int warps = (g_tune.warps >= 1 && g_tune.warps <= MAX_WARPS) ? g_tune.warps : (items >= 32 * fdtd::sm_count() ? 8 : (items >= 8 * fdtd::sm_count() ? 4 : 2));
    while (warps > 1 && (size_t)warps * per_warp > 200 * 1024) --warps;
    const size_t smem = (size_t)warps * per_warp;
    if (smem > 48 * 1024) {
        // per instantiation and device: largest dynamic smem opted in so far (shared by every host thread)
        static size_t configured[64] = {0};
        static std::mutex guard;
        std::lock_guard<std::mutex> lock(guard);
        int dev = 0;
        FDTD_CUDA(cudaGetDevice(&dev));
        size_t &opted = configured[dev >= 0 && dev < 64 ? dev : 0];
        if (smem > opted || dev >= 64) {
            FDTD_CUDA(cudaFuncSetAttribute(k_march<real, V, T, MODE, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            opted = smem;
        }
    }
    const int grid = (items + warps - 1) / warps;
    k_march<real, V, T, MODE, FAST><<<grid, warps * 32, smem, st>>>(mp, all_careful);
    FDTD_LAUNCH_CHECK("k_march");
    return FDTD_OK;
}

// classify strips and chunks on the host (classify_pass: the conditions the kernels rely on) and launch the two kernels
template <typename real, int V, int T, int MODE>
int launch_march_k(MarchParams<real> &mp, cudaStream_t st) {
    const PassCounts pc = classify_pass(mp, V, T);
    // the careful (edge) warps: the register-shifting kernel of this file or, on request (fdtd2d_tune2 deep = 2), the
    // shared-memory-resident one of fd2d_deep.cu
    auto launch_careful = [&](int items, int all, cudaStream_t where) -> int {
        if constexpr (sizeof(real) == 4 && (V == 4 || V == 2) && (MODE & 2) == 0) {
            if (g_tune.deep == 2) return launch_careful2(mp, T, (MODE & 1) != 0, items, all, where, V);
        }
        return launch_one<real, V, T, MODE, false>(mp, items, all, where);
    };
    if (pc.all_careful) return launch_careful(pc.n_careful, 1, st);
    const int n_fast = pc.n_fast, n_careful = pc.n_careful;
    // The careful kernel is small (edges only) and would run alone at a fraction of a wave: fork it onto a side
    // stream so the interior kernel backfills the SMs it leaves idle, and join before the next pass.
    SideStream *side = (n_careful > 0 && n_fast > 0 && g_tune.serial >= 2) ? side_stream(st) : nullptr;
    // interior warps: one launch, or -- lossy problem with a lossless-outside promise -- two over the same index space
    // (the lossy kernel keeps the warps that meet the box, the lossless kernel the others)
    auto launch_interior = [&]() -> int {
        if constexpr (MODE == 1) {
            if (mp.split_lossless) {
                int rc2 = launch_one<real, V, T, 0, true>(mp, n_fast, 0, st);
                if (rc2 != FDTD_OK) return rc2;
            }
        }
        return launch_one<real, V, T, MODE, true>(mp, n_fast, 0, st);
    };
    if (side == nullptr) {
        int rc = launch_careful(n_careful, 0, st);
        if (rc != FDTD_OK) return rc;
        return launch_interior();
    }
    FDTD_CUDA(cudaEventRecord(side->fork, st));
    FDTD_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
    int rc = launch_careful(n_careful, 0, side->stream);
    if (rc != FDTD_OK) return rc;
    FDTD_CUDA(cudaEventRecord(side->join, side->stream));
    rc = launch_interior();
    if (rc != FDTD_OK) return rc;
    FDTD_CUDA(cudaStreamWaitEvent(st, side->join, 0));
    return FDTD_OK;
}

template <typename real, int V, int T>
int launch_march(MarchParams<real> &mp, bool lossy, cudaStream_t st) {
    if (mp.nf > 0) {           // running DFT fused into the pass: narrow vectors, shallow blocking (6 more registers per cell)
        if constexpr (V <= 2 && T <= 4) {
            return lossy ? launch_march_k<real, V, T, 3>(mp, st) : launch_march_k<real, V, T, 2>(mp, st);
        } else {
            fdtd::set_error("no fused-DFT kernel for vector width %d / depth %d", V, T);
            return FDTD_EUNSUPPORTED;
        }
    }
    return lossy ? launch_march_k<real, V, T, 1>(mp, st) : launch_march_k<real, V, T, 0>(mp, st);
}

template <typename real, int V>
int launch_march_T(int T, MarchParams<real> &mp, bool lossy, cudaStream_t st) {
#ifdef FDTD_DEV_ONLY        // development builds: one instantiation family (seconds instead of minutes of ptxas)
    if constexpr (sizeof(real) == 4 && V == 4) {
        if (T == 6) return launch_march<real, V, 6>(mp, lossy, st);
        if (T == 8) return launch_march<real, V, 8>(mp, lossy, st);
    }
    fdtd::set_error("development build: float, V=4, T in {6, 8} only");
    return FDTD_EUNSUPPORTED;
#else
    switch (T) {
        case 1: return launch_march<real, V, 1>(mp, lossy, st);
        case 2: return launch_march<real, V, 2>(mp, lossy, st);
        case 3: return launch_march<real, V, 3>(mp, lossy, st);
        case 4: return launch_march<real, V, 4>(mp, lossy, st);
        case 6: return launch_march<real, V, 6>(mp, lossy, st);
        case 8: return launch_march<real, V, 8>(mp, lossy, st);
        default: fdtd::set_error("unsupported time-block depth %d for vector width %d", T, V); return FDTD_EUNSUPPORTED;
    }
#endif
}

// CUDA loads kernels lazily: the first launch of every instantiation pays a module load of several ms.
// preload() resolves all instantiations one problem can reach (every depth, both kernels) ahead of time.
template <typename real, int V, int T>
void touch_T(bool lossy) {
    cudaFuncAttributes a;
    if (lossy) {
        cudaFuncGetAttributes(&a, k_march<real, V, T, 1, true>);
        cudaFuncGetAttributes(&a, k_march<real, V, T, 1, false>);
        cudaFuncGetAttributes(&a, k_march<real, V, T, 0, true>);     // lossless-outside split
    } else {
        cudaFuncGetAttributes(&a, k_march<real, V, T, 0, true>);
        cudaFuncGetAttributes(&a, k_march<real, V, T, 0, false>);
    }
}
template <typename real, int V, int T>
void touch_dft(bool lossy) {
    cudaFuncAttributes a;
    if (lossy) {
        cudaFuncGetAttributes(&a, k_march<real, V, T, 3, true>);
        cudaFuncGetAttributes(&a, k_march<real, V, T, 3, false>);
    } else {
        cudaFuncGetAttributes(&a, k_march<real, V, T, 2, true>);
        cudaFuncGetAttributes(&a, k_march<real, V, T, 2, false>);
    }
}

template <typename real, int V>
void touch_V(bool lossy) {
    touch_T<real, V, 1>(lossy); touch_T<real, V, 2>(lossy); touch_T<real, V, 3>(lossy);
    touch_T<real, V, 4>(lossy); touch_T<real, V, 6>(lossy);
    touch_T<real, V, 8>(lossy);
}

// vector width: widest V dividing ny (rows start V-aligned; the strip halo is rounded up to a multiple of V)
template <typename real> int pick_v(int ny);
template <> int pick_v<float>(int ny) { return ny % 4 == 0 ? 4 : (ny % 2 == 0 ? 2 : 1); }
template <> int pick_v<double>(int ny) { return ny % 2 == 0 ? 2 : 1; }

// Launch shape by grid size (cells stored on this device), from the sweeps in profiles/r1_sweep_grid_sizes_3_2.txt:
// small grids live in L2 and need many short warps (narrow vectors, shallow blocking, 16-row chunks); big grids
// are HBM-bound and want 4-wide vectors, 6 steps per pass and 128-row chunks.
struct Plan {
    int V, T, chunk;
    bool deep;       // the deep passes of fd2d_deep.cu (depth 8 and 12: float, 4-wide vectors, no fused DFT) are available
    long cells;      // cells the call produces
};

// rows per chunk of a pass.  The deep (warp-chain) passes recompute 2T rows per chunk and, with the short edge chunks,
// pay nothing extra at the edges for tall chunks: 384 rows at 32768^2 (965 Gcell/s; 256: 947, 128: 891), 256 at
// 16384^2 (832; 128: 798), 128 at 8192^2 (640; 64: 604, 256: 605) -- profiles/r2_chunk_rows_by_grid.txt.  The
// register-pipeline passes keep their round-1 heights (depth 6 at 32768^2: 128 rows 781, 192: 763, 256: 751).
inline int pass_chunk_rows(const Plan &plan, int T, bool deep_pass) {
    if (g_tune.chunk_rows > 0) return g_tune.chunk_rows;
    if (deep_pass) return plan.cells >= 500000000L ? 384 : (plan.cells >= 100000000L ? 256 : 128);
    return max(plan.chunk, 4 * T);                  // at most ~50 % warm-up/drain recompute on shallow grids
}

template <typename real>
Plan choose_plan(const fdtd2d_problem *q) {
    // sized by the rows THIS call produces (a slab, or one row block of a streamed run), not by the allocation
    const long cells = (long)(q->row_hi - q->row_lo) * q->ny;
    const bool lossy = (q->flags & FDTD_LOSSY) != 0;
    const int ny = q->ny;
    Plan p;
    if (sizeof(real) == 4) {                            // profiles/r1_v18_sweep_grid_sizes.txt
        if (cells < 6000000L)        p = {2, 4, 16, false, cells};
        else if (cells < 24000000L)  p = {2, 6, 64, false, cells};
        else if (cells < 100000000L) p = {lossy ? 2 : 4, 6, 64, false, cells};
        else                         p = {lossy ? 2 : 4, 6, 128, false, cells};   // 7 row sets x 4 columns x 9 lossy fields spill
    } else {
        if (cells < 1500000L)        p = {1, 4, 16, false, cells};           // profiles/r1_sweep_fp64.txt
        else if (cells < 12000000L)  p = {2, 4, 16, false, cells};
        else if (cells < 150000000L) p = {2, 6, 64, false, cells};
        else                         p = {2, 6, 128, false, cells};
    }
    while (p.V > 1 && ny % p.V != 0) p.V >>= 1;
    // vector width of the call (the depth-dependent float64 exception is applied per pass)
    int V = g_tune.force_v ? g_tune.force_v : p.V;
    if (q->nf > 0 && V == 4) V = 2;                 // fused-DFT kernels: vector width <= 2
    if (q->nf > 0 && sizeof(real) == 8) V = 1;      // ... and 1 in float64 (register row sets)
    if (ny % V != 0 || (sizeof(real) == 8 && V == 4)) V = 1;
    p.V = V;
    p.deep = sizeof(real) == 4 && V == 4 && q->nf == 0 && g_tune.deep != 0 && deep_supported(12, lossy);
    return p;
}

// Depth of the next pass: at most `left` steps and `tblock`, rounded down to an instantiated depth: 1, 2, 3, 4, 6, 8, 12
// (8 and 12 where the deep passes apply).  tblock = 0: the library's choice.  Where the deep passes apply that is the
// split of the remaining steps into passes of least total cost: every pass moves the state through HBM once, so a
// shallow pass costs about as much as a depth-6 one (HBM-bound: 8.25 ms at 32768^2), a depth-8 pass of the warp-chain
// kernel 8.9 ms and a depth-12 pass 14.9 ms (profiles/r2_chunk_rows_by_grid.txt) -- e.g. 20 steps = 12 + 8 (23.8 ms;
// 8 + 6 + 6 = 25.4, round 1's 6 + 6 + 6 + 2 = 34.2), 96 steps = 12 x 8.  Depth 12 is measured at 32768^2 only and is
// offered to the split from 500 M cells up.
inline int next_depth(const Plan &plan, int tblock, int left, int nf) {
    if (tblock <= 0 && plan.deep && nf == 0) {
        // cost of a pass by depth, in units of a depth-6 pass; least-cost split by dynamic programming over `left`
        static const int depths[] = {12, 8, 6, 4, 3, 2, 1};
        static const double cost[] = {1.81, 1.09, 1.00, 0.97, 0.96, 0.95, 0.94};
        const int k0 = plan.cells >= 500000000L ? 0 : 1;
        constexpr int HORIZON = 48;                     // beyond this many steps the split starts with a depth-8 pass anyway
        if (left > HORIZON) return 8;
        double best[HORIZON + 1];
        int first[HORIZON + 1];
        best[0] = 0.0; first[0] = 0;
        for (int n = 1; n <= left; ++n) {
            best[n] = 1e30; first[n] = 1;
            for (int k = k0; k < 7; ++k) {
                if (depths[k] > n) continue;
                const double c = cost[k] + best[n - depths[k]];
                if (c < best[n] - 1e-9) { best[n] = c; first[n] = depths[k]; }
            }
        }
        return first[left];
    }
    int T = min(tblock > 0 ? tblock : plan.T, left);
    if (nf > 0) T = min(T, 4);                      // fused-DFT kernels: depth <= 4
    if (T >= 12 && plan.deep) return 12;
    if (T > 8) T = 8;
    if (T == 5 || T == 7) --T;
    return T;
}

template <typename real>
int advance(const fdtd2d_problem *q, int cur, int nsteps, const double *src, int tblock, cudaStream_t st,
            int *cur_out) {
    const Plan plan = choose_plan<real>(q);
    const bool lossy = (q->flags & FDTD_LOSSY) != 0, tfsf = (q->flags & FDTD_TFSF) != 0;
    int done = 0;
    while (done < nsteps) {
        const int T = next_depth(plan, tblock, nsteps - done, q->nf);
        if (q->halo > 0 && T < nsteps) {
            // the handshake orders whole calls: the last pass pushes into the array set that the neighbour's earlier
            // passes of the same call would still be reading
            fdtd::set_error("fdtd2d_advance: the fused halo exchange takes one pass per call (%d steps asked, pass depth %d)", nsteps, T);
            return FDTD_EINVAL;
        }
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
        // ez is an output only (every pass recomputes it from dz): intermediate passes need not store it.
        // With a lossy medium it cannot be rebuilt afterwards (iz has moved on), so it is always stored.
        mp.write_ez = lossy || (rem == 0 && !(q->flags & FDTD_LAZY_EZ));
        mp.ident_row_lo = q->ident_row_lo; mp.ident_row_hi = q->ident_row_hi;
        mp.ident_col_lo = q->ident_col_lo; mp.ident_col_hi = q->ident_col_hi;
        mp.ezi_hist = (const real *)q->ezi_hist; mp.hxi_hist = (const real *)q->hxi_hist;
        mp.src_i = tfsf ? -1 : q->src_i; mp.src_j = q->src_j; mp.src_hard = q->src_hard;
        for (int s = 0; s < TMAX; ++s) mp.src[s] = (src && s < T) ? src[done + s] : 0.0;
        mp.negzero2 = 0x8000000080000000ull;
        mp.spin_ns = g_tune.spin_ns;
        {   // fused halo exchange: wait on the first pass of the call, push + announce on the last
            const bool on = q->halo > 0;
            void *const *up = q->peer_up[cur ^ 1];
            void *const *dn = q->peer_dn[cur ^ 1];
            mp.push = on && rem == 0;
            mp.halo_on = on; mp.own_lo = q->row_lo; mp.own_hi = q->row_hi;
            mp.wait_flags = on && done == 0;
            mp.signal = on && rem == 0;
            mp.push_up_end = q->row_lo + q->halo;
            mp.push_dn_begin = q->row_hi - q->halo;
            mp.up_dz = on ? (real *)up[FDTD2D_DZ] : nullptr;   mp.up_hx = on ? (real *)up[FDTD2D_HX] : nullptr;
            mp.up_hy = on ? (real *)up[FDTD2D_HY] : nullptr;   mp.up_ihx = on ? (real *)up[FDTD2D_IHX] : nullptr;
            mp.up_ihy = on ? (real *)up[FDTD2D_IHY] : nullptr; mp.up_iz = on ? (real *)up[FDTD2D_IZ] : nullptr;
            mp.dn_dz = on ? (real *)dn[FDTD2D_DZ] : nullptr;   mp.dn_hx = on ? (real *)dn[FDTD2D_HX] : nullptr;
            mp.dn_hy = on ? (real *)dn[FDTD2D_HY] : nullptr;   mp.dn_ihx = on ? (real *)dn[FDTD2D_IHX] : nullptr;
            mp.dn_ihy = on ? (real *)dn[FDTD2D_IHY] : nullptr; mp.dn_iz = on ? (real *)dn[FDTD2D_IZ] : nullptr;
            mp.up_shift = (long long)(q->row_base - q->peer_up_base) * q->ny;
            mp.dn_shift = (long long)(q->row_base - q->peer_dn_base) * q->ny;
            mp.sync_local = (unsigned long long *)q->sync_local;
            mp.flag_at_up = (on && up[FDTD2D_DZ]) ? (unsigned long long *)q->sync_up + 1 : nullptr;   // I am its "down"
            mp.flag_at_dn = (on && dn[FDTD2D_DZ]) ? (unsigned long long *)q->sync_dn + 0 : nullptr;   // I am its "up"
            mp.epoch = q->epoch;
            mp.total_warps = 0;
        }
        mp.lz_row_lo = q->lossy_row_lo; mp.lz_row_hi = q->lossy_row_hi;
        mp.lz_col_lo = q->lossy_col_lo; mp.lz_col_hi = q->lossy_col_hi;
        mp.split_lossless = lossy && q->nf == 0 && g_tune.split != 0 && q->lossy_row_hi > q->lossy_row_lo && q->lossy_col_hi > q->lossy_col_lo;
        mp.nf = q->nf;
        mp.r_pt = (real *)q->ft.r_pt; mp.i_pt = (real *)q->ft.i_pt;
        mp.dft_plane = (long long)q->rows_alloc * q->ny;
        for (int s = 0; s < TMAX; ++s)
            for (int f = 0; f < NFMAX; ++f) {
                const bool on = q->nf > 0 && s < T && f < q->nf;
                mp.dft_c[s][f] = on ? q->dft_cos[(size_t)(done + s) * q->nf + f] : 0.0;
                mp.dft_s[s][f] = on ? q->dft_sin[(size_t)(done + s) * q->nf + f] : 0.0;
            }

        int V = plan.V;
        if (T == 8 && sizeof(real) == 8) V = 1;         // 9 register sets of 2 doubles do not fit
        // deep pass (fd2d_deep.cu): accumulators resident in shared memory, depth 12 -- and depth 8 in place of the
        // register-pipeline kernel with its naz ring
        const bool deep = plan.deep && (T == 12 || T == 8);
        const int halo = ((T + V - 1) / V) * V;
        const int use = 32 * V - 2 * halo;
        mp.nstrips = (q->ny + use - 1) / use;
        const int rows = mp.out_hi - mp.out_lo;
        int chunk = pass_chunk_rows(plan, T, deep);
        chunk = max(1, min(chunk, rows));
        mp.chunk_rows = chunk;
        mp.nchunks = (rows + chunk - 1) / chunk;
        mp.first_rows = mp.last_rows = 0;
        mp.col_fast = mp.row_fast = 0;
        if (g_tune.edge_chunks && rows >= 3 * chunk) {
            // Short edge chunks: F = the fewest rows (a multiple of 8) after which a chunk starts clear of everything
            // that makes the rows above it special -- the same conditions classify_pass applies to lo = i0 - T - 1 --
            // and L likewise for hi = i1 + 2T + RING + 2 below.  No edge condition on a side: that chunk stays whole.
            const int margin_lo = T + 1, margin_hi = 2 * T + RING + 2;
            int top = max(max(1, mp.in_lo), mp.ident_row_lo);          // lo must not fall below this row
            int bot = min(min(mp.nx - 1, mp.in_hi), mp.ident_row_hi);  // hi must not exceed this row
            int f_req = top + margin_lo - mp.out_lo, l_req = mp.out_hi + margin_hi - bot;
            if (q->halo > 0) {
                f_req = max(f_req, max(q->row_lo + margin_lo, q->row_lo + q->halo) - mp.out_lo);
                l_req = max(l_req, mp.out_hi - min(q->row_hi - margin_hi, q->row_hi - q->halo));
            }
            int F = f_req <= 0 ? chunk : min(chunk, (f_req + 7) / 8 * 8);
            int L = l_req <= 0 ? chunk : min(chunk, (l_req + 7) / 8 * 8);
            if ((F < chunk || L < chunk) && rows - F - L >= 1) {
                mp.first_rows = F;
                mp.last_rows = L;
                mp.nchunks = 2 + (rows - F - L + chunk - 1) / chunk;
            }
        }

        if (tfsf && !(q->flags & FDTD_INCIDENT_READY)) {
            SrcTable tab;
            for (int s = 0; s < TMAX; ++s) tab.v[s] = mp.src[s];
            tab.nf = (mp.nf > 0 && q->ft.r_in && q->ft.i_in) ? mp.nf : 0;
            for (int s = 0; s < TMAX; ++s)
                for (int f = 0; f < NFMAX; ++f) { tab.c[s][f] = mp.dft_c[s][f]; tab.s[s][f] = mp.dft_s[s][f]; }
            k_incident_line<real><<<1, 1024, 0, st>>>(q->ny, q->npml, T, (real *)q->ezi, (real *)q->hxi, (real *)q->bc,
                                                      (real *)q->ezi_hist, (real *)q->hxi_hist, (real *)q->ft.r_in,
                                                      (real *)q->ft.i_in, tab);
            FDTD_LAUNCH_CHECK("k_incident_line");
        }
        int rc;
        if constexpr (sizeof(real) == 4) {
            if (deep) rc = launch_march_deep(mp, T, lossy, st);
            else if (V == 4) rc = launch_march_T<real, 4>(T, mp, lossy, st);
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

int fdtd2d_check_identity(const fdtd2d_problem *q, long long *violations) {
    FDTD_REQUIRE(q && violations, "fdtd2d_check_identity: null argument");
    FDTD_REQUIRE(q->dtype == FDTD_F32 || q->dtype == FDTD_F64, "fdtd2d_check_identity: unknown dtype %d", q->dtype);
    FDTD_REQUIRE(q->ident_row_lo >= 0 && q->ident_row_hi <= q->nx && q->ident_col_lo >= 0 && q->ident_col_hi <= q->ny,
                 "fdtd2d_check_identity: identity ranges outside the grid");
    unsigned long long *bad = nullptr, host = 0;
    FDTD_CUDA(cudaMalloc(&bad, sizeof(*bad)));
    FDTD_CUDA(cudaMemset(bad, 0, sizeof(*bad)));
    const int nr = q->ident_row_hi - q->ident_row_lo, nc = q->ident_col_hi - q->ident_col_lo;
    if (q->dtype == FDTD_F32) {
        using R = float;
        if (nr > 0) k_check_identity<R><<<(nr + 255) / 256, 256>>>((const R *)q->pml.fx1, (const R *)q->pml.fx2, (const R *)q->pml.fx3, (const R *)q->pml.gx2, (const R *)q->pml.gx3, q->ident_row_lo, q->ident_row_hi, bad);
        if (nc > 0) k_check_identity<R><<<(nc + 255) / 256, 256>>>((const R *)q->pml.fy1, (const R *)q->pml.fy2, (const R *)q->pml.fy3, (const R *)q->pml.gy2, (const R *)q->pml.gy3, q->ident_col_lo, q->ident_col_hi, bad);
    } else {
        using R = double;
        if (nr > 0) k_check_identity<R><<<(nr + 255) / 256, 256>>>((const R *)q->pml.fx1, (const R *)q->pml.fx2, (const R *)q->pml.fx3, (const R *)q->pml.gx2, (const R *)q->pml.gx3, q->ident_row_lo, q->ident_row_hi, bad);
        if (nc > 0) k_check_identity<R><<<(nc + 255) / 256, 256>>>((const R *)q->pml.fy1, (const R *)q->pml.fy2, (const R *)q->pml.fy3, (const R *)q->pml.gy2, (const R *)q->pml.gy3, q->ident_col_lo, q->ident_col_hi, bad);
    }
    FDTD_LAUNCH_CHECK("k_check_identity");
    FDTD_CUDA(cudaMemcpy(&host, bad, sizeof(host), cudaMemcpyDeviceToHost));
    FDTD_CUDA(cudaFree(bad));
    *violations = (long long)host;
    return FDTD_OK;
}

int fdtd2d_check_lossless_outside(const fdtd2d_problem *q, long long *violations) {
    FDTD_REQUIRE(q && violations, "fdtd2d_check_lossless_outside: null argument");
    FDTD_REQUIRE(q->dtype == FDTD_F32 || q->dtype == FDTD_F64, "fdtd2d_check_lossless_outside: unknown dtype %d", q->dtype);
    FDTD_REQUIRE((q->flags & FDTD_LOSSY) && q->md.nbz && q->state[0][FDTD2D_IZ] && q->state[1][FDTD2D_IZ],
                 "fdtd2d_check_lossless_outside: a lossy problem with nbz and both iz arrays");
    FDTD_REQUIRE(q->rows_alloc >= 1 && q->ny >= 1, "fdtd2d_check_lossless_outside: bad sizes");
    unsigned long long *bad = nullptr, host = 0;
    FDTD_CUDA(cudaMalloc(&bad, sizeof(*bad)));
    FDTD_CUDA(cudaMemset(bad, 0, sizeof(*bad)));
    const dim3 grid((q->ny + 255) / 256, min(q->rows_alloc, 65535));
    if (q->dtype == FDTD_F32)
        k_check_lossless_outside<float><<<grid, 256>>>((const float *)q->md.nbz, (const float *)q->state[0][FDTD2D_IZ], (const float *)q->state[1][FDTD2D_IZ],
                                                       q->ny, q->row_base, q->rows_alloc, q->lossy_row_lo, q->lossy_row_hi, q->lossy_col_lo, q->lossy_col_hi, bad);
    else
        k_check_lossless_outside<double><<<grid, 256>>>((const double *)q->md.nbz, (const double *)q->state[0][FDTD2D_IZ], (const double *)q->state[1][FDTD2D_IZ],
                                                        q->ny, q->row_base, q->rows_alloc, q->lossy_row_lo, q->lossy_row_hi, q->lossy_col_lo, q->lossy_col_hi, bad);
    FDTD_LAUNCH_CHECK("k_check_lossless_outside");
    FDTD_CUDA(cudaMemcpy(&host, bad, sizeof(host), cudaMemcpyDeviceToHost));
    FDTD_CUDA(cudaFree(bad));
    *violations = (long long)host;
    return FDTD_OK;
}

int fdtd2d_preload(int dtype, int ny, int lossy) {
    (void)ny;
#ifdef FDTD_DEV_ONLY
    (void)dtype; (void)lossy;
    return FDTD_OK;
#else
    if (lossy & 2) {           // bit 1: the fused-DFT kernels as well
        const bool l = (lossy & 1) != 0;
        if (dtype == FDTD_F32) {
            touch_dft<float, 2, 1>(l); touch_dft<float, 2, 2>(l); touch_dft<float, 2, 3>(l); touch_dft<float, 2, 4>(l);
            touch_dft<float, 1, 1>(l); touch_dft<float, 1, 2>(l); touch_dft<float, 1, 3>(l); touch_dft<float, 1, 4>(l);
        } else if (dtype == FDTD_F64) {
            touch_dft<double, 2, 1>(l); touch_dft<double, 2, 2>(l); touch_dft<double, 2, 3>(l); touch_dft<double, 2, 4>(l);
            touch_dft<double, 1, 1>(l); touch_dft<double, 1, 2>(l); touch_dft<double, 1, 3>(l); touch_dft<double, 1, 4>(l);
        }
    }
    lossy &= 1;
    if (dtype == FDTD_F32) {
        preload_deep(lossy != 0);
        touch_V<float, 4>(lossy != 0);
        touch_V<float, 2>(lossy != 0);
        touch_V<float, 1>(lossy != 0);
    } else if (dtype == FDTD_F64) {
        touch_V<double, 2>(lossy != 0);
        touch_V<double, 1>(lossy != 0);
    } else {
        fdtd::set_error("fdtd2d_preload: unknown dtype %d", dtype);
        return FDTD_EINVAL;
    }
    FDTD_LAUNCH_CHECK("fdtd2d_preload");
    return FDTD_OK;
#endif
}

int fdtd2d_max_tblock(int dtype, int ny) {
    if (dtype == FDTD_F32) return (ny % 4 == 0 && deep_supported(12, false)) ? 12 : 8;
    return dtype == FDTD_F64 ? 8 : 0;
}

// tuning / test hook (not part of the reference-facing surface): force the vector width and rows per chunk
int fdtd2d_tune(int force_v, int chunk_rows, int warps_per_cta, int ring_depth, int force_careful) {
    g_tune.force_v = force_v;
    g_tune.chunk_rows = chunk_rows;
    g_tune.warps = warps_per_cta;
    g_tune.serial = (ring_depth == 1) ? 1 : (ring_depth == 3 ? 3 : 2);     // (slot reused) 1 = serialise the edge and interior kernels, 3 = edge kernel backfills
    g_tune.careful = force_careful & 1;            // bit 0: every warp through the careful kernel
    g_tune.split = (force_careful & 2) ? 0 : 1;    // bit 1: ignore the lossless-outside promise
    return FDTD_OK;
}

int fdtd2d_tune2(int key, long long value) {
    switch (key) {
        case FDTD_TUNE_DEEP:
            FDTD_REQUIRE(value >= 0 && value <= 2, "fdtd2d_tune2: deep must be 0, 1 or 2");
            g_tune.deep = (int)value;
            return FDTD_OK;
        case FDTD_TUNE_HALO_WAIT_MS:
            FDTD_REQUIRE(value >= 0, "fdtd2d_tune2: negative wait");
            g_tune.spin_ns = value == 0 ? HALO_SPIN_NS : (unsigned long long)value * 1000000ull;
            return FDTD_OK;
        case FDTD_TUNE_VARIANT:
            g_tune.variant = (int)value;
            return FDTD_OK;
        case FDTD_TUNE_EDGE_CHUNKS:
            g_tune.edge_chunks = value != 0;
            return FDTD_OK;
        case FDTD_TUNE_COL_FAST:
            g_tune.col_fast = (value & 1) != 0;            // bit 0: column variant, bit 1: row variant
            g_tune.row_fast = (value & 2) != 0;
            return FDTD_OK;
        default:
            fdtd::set_error("fdtd2d_tune2: unknown key %d", key);
            return FDTD_EINVAL;
    }
}

int fdtd2d_plan(const fdtd2d_problem *q, int nsteps, int tblock, int *depths, int cap, int *vector_width, int *chunk_rows) {
    FDTD_REQUIRE(q != nullptr, "fdtd2d_plan: null problem");
    FDTD_REQUIRE(q->dtype == FDTD_F32 || q->dtype == FDTD_F64, "fdtd2d_plan: unknown dtype %d", q->dtype);
    FDTD_REQUIRE(nsteps >= 0 && tblock >= 0 && tblock <= TMAX, "fdtd2d_plan: nsteps %d / tblock %d out of range", nsteps, tblock);
    FDTD_REQUIRE(q->row_hi > q->row_lo && q->ny >= 2, "fdtd2d_plan: empty problem");
    const Plan plan = q->dtype == FDTD_F32 ? choose_plan<float>(q) : choose_plan<double>(q);
    int n = 0, first = 0;
    for (int left = nsteps; left > 0; ++n) {
        const int T = next_depth(plan, tblock, left, q->nf);
        if (n == 0) first = T;
        if (depths != nullptr && n < cap) depths[n] = T;
        left -= T;
    }
    if (vector_width) *vector_width = (q->dtype == FDTD_F64 && first == 8) ? 1 : plan.V;
    if (chunk_rows) *chunk_rows = pass_chunk_rows(plan, max(first, 1), plan.deep && (first == 8 || first == 12));
    return n;
}

int fdtd2d_incident_line(const fdtd2d_problem *q, int nsteps, const double *src, void *ezi_hist, void *hxi_hist, void *stream) {
    FDTD_REQUIRE(q && src && ezi_hist && hxi_hist, "fdtd2d_incident_line: null argument");
    FDTD_REQUIRE(q->dtype == FDTD_F32 || q->dtype == FDTD_F64, "fdtd2d_incident_line: unknown dtype %d", q->dtype);
    FDTD_REQUIRE((q->flags & FDTD_TFSF) && q->ezi && q->hxi && q->bc, "fdtd2d_incident_line: a TFSF problem with its incident line");
    FDTD_REQUIRE(nsteps >= 1 && nsteps <= TMAX, "fdtd2d_incident_line: %d steps outside [1, %d]", nsteps, TMAX);
    FDTD_REQUIRE(q->npml >= 2 && 2 * q->npml <= q->ny && q->ny >= 8, "fdtd2d_incident_line: bad TFSF geometry");
    SrcTable tab;
    for (int s = 0; s < TMAX; ++s) tab.v[s] = s < nsteps ? src[s] : 0.0;
    tab.nf = 0;
    for (int s = 0; s < TMAX; ++s)
        for (int f = 0; f < NFMAX; ++f) tab.c[s][f] = tab.s[s][f] = 0.0;
    if (q->dtype == FDTD_F32)
        k_incident_line<float><<<1, 1024, 0, fdtd::as_stream(stream)>>>(q->ny, q->npml, nsteps, (float *)q->ezi, (float *)q->hxi, (float *)q->bc,
                                                                        (float *)ezi_hist, (float *)hxi_hist, nullptr, nullptr, tab);
    else
        k_incident_line<double><<<1, 1024, 0, fdtd::as_stream(stream)>>>(q->ny, q->npml, nsteps, (double *)q->ezi, (double *)q->hxi, (double *)q->bc,
                                                                         (double *)ezi_hist, (double *)hxi_hist, nullptr, nullptr, tab);
    FDTD_LAUNCH_CHECK("k_incident_line");
    return FDTD_OK;
}

int fdtd2d_halo_status(const fdtd2d_problem *q, unsigned long long *word) {
    FDTD_REQUIRE(q && word && q->sync_local, "fdtd2d_halo_status: null argument / no fused halo exchange");
    FDTD_CUDA(cudaMemcpy(word, (const unsigned long long *)q->sync_local + 4, sizeof(*word), cudaMemcpyDeviceToHost));
    if (*word != 0)
        fdtd::set_error("fused halo exchange: the pass of epoch %llu gave up waiting for its %s neighbour (a rank died or skipped a call); "
                        "ghost rows are stale", *word >> 2, (*word & 3) == 1 ? "upper" : "lower");
    return FDTD_OK;
}

int fdtd2d_advance(const fdtd2d_problem *q, int cur, int nsteps, const double *src, int tblock, void *stream,
                   int *cur_out) {
    FDTD_REQUIRE(q && cur_out, "fdtd2d_advance: null problem / cur_out");
    FDTD_REQUIRE(cur == 0 || cur == 1, "fdtd2d_advance: cur must be 0 or 1");
    FDTD_REQUIRE(q->nx >= 2 && q->ny >= 2, "fdtd2d_advance: grid %dx%d too small", q->nx, q->ny);
    FDTD_REQUIRE(tblock >= 0 && tblock <= TMAX, "fdtd2d_advance: tblock %d outside [0, %d] (0 = choose by grid size)", tblock, TMAX);
    FDTD_REQUIRE(nsteps >= 0, "fdtd2d_advance: nsteps < 0");
    FDTD_REQUIRE(q->row_lo >= 0 && q->row_hi <= q->nx && q->row_lo < q->row_hi, "fdtd2d_advance: bad owned rows [%d,%d)", q->row_lo, q->row_hi);
    FDTD_REQUIRE(q->row_base <= q->row_lo && q->row_base + q->rows_alloc >= q->row_hi, "fdtd2d_advance: owned rows outside the stored rows");
    if (!(q->flags & FDTD_GHOST_DECAY)) {   // every row within nsteps of the owned range (clipped to the grid) must be stored
        const int need_lo = q->row_lo - nsteps > 0 ? q->row_lo - nsteps : 0;
        const int need_hi = q->row_hi + nsteps < q->nx ? q->row_hi + nsteps : q->nx;
        FDTD_REQUIRE(q->row_base <= need_lo && q->row_base + q->rows_alloc >= need_hi,
                     "fdtd2d_advance: %d steps need ghost rows [%d,%d) but only [%d,%d) are stored", nsteps, need_lo,
                     need_hi, q->row_base, q->row_base + q->rows_alloc);
    }
    if (q->halo > 0) {
        FDTD_REQUIRE(q->halo <= q->row_hi - q->row_lo && q->sync_local && q->epoch >= 1, "fdtd2d_advance: fused halo exchange needs halo <= owned rows, sync_local and epoch >= 1");
        FDTD_REQUIRE(!q->peer_up[0][FDTD2D_DZ] == !q->sync_up && !q->peer_dn[0][FDTD2D_DZ] == !q->sync_dn, "fdtd2d_advance: a peer needs both its arrays and its sync words");
        FDTD_REQUIRE(nsteps <= q->halo, "fdtd2d_advance: %d steps but only %d halo rows are exchanged", nsteps, q->halo);
    }
    FDTD_REQUIRE(q->nf >= 0 && q->nf <= NFMAX, "fdtd2d_advance: nf=%d outside [0, %d] (use fdtd2d_fourier per step beyond)", q->nf, NFMAX);
    FDTD_REQUIRE(q->nf == 0 || (q->ft.r_pt && q->ft.i_pt && q->dft_cos && q->dft_sin), "fdtd2d_advance: running DFT needs r_pt, i_pt and the phase tables");
    FDTD_REQUIRE(q->ident_row_lo >= 0 && q->ident_row_hi <= q->nx && q->ident_col_lo >= 0 && q->ident_col_hi <= q->ny,
                 "fdtd2d_advance: identity ranges outside the grid");
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
