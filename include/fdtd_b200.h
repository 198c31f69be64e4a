/* fdtd_b200.h -- C ABI of libfdtd_b200.so: B200 (sm_100a) kernels for the Yee time-stepping hot path of
 * dsarvan/simulation (fd1d/: Ex/Hy; fd2d/: TM Dz/Ez/Hx/Hy + PML + TFSF + lossy medium).
 *
 * The reference has no FFI; its operator interface is the module-level step-function protocol shared by
 * all of its language variants, and -- for a Python host driving GPU kernels -- the PyCUDA idiom
 *     fn = SourceModule(src).get_function("dfield"); fn(t, nx, ny, pml, ezi, dz, hx, hy, grid=, block=)
 * (reference fd2d/pycuda/test_3_3.py:247-270, fd2d/cuda/test_3_4.cu:20-36,333-343).  This header keeps those
 * function names, argument order and in-place/caller-owns-everything semantics, on raw device pointers.
 *
 * Conventions
 *   - every entry point returns 0 on success, a negative FDTD_E* code on failure; fdtd_last_error()
 *     returns a thread-local message.  The reference has no error reporting at all.
 *   - `dtype` is FDTD_F32 or FDTD_F64; every array of one call has that element type.
 *   - all array arguments are DEVICE pointers (16-byte aligned); 2D arrays are C-order (nx rows, ny
 *     columns, j fastest), exactly the reference layout n = i*ny + j (fd2d/cuda/test_3_4.cu:80).
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Calls are asynchronous.
 *   - source waveforms are evaluated on the HOST in float64 (the reference hard-codes them inside
 *     dfield/exfield/dxfield): a hard source overwrites, a soft source is added in float64 and rounded
 *     once -- the numpy semantics of `ex[1] += np.float64`.
 *   - arithmetic is evaluated in the reference's left-to-right order with no FMA contraction, so results
 *     are bit-identical to the reference numpy programs.
 */
#ifndef FDTD_B200_H
#define FDTD_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FDTD_F32 0
#define FDTD_F64 1

#define FDTD_MAX_FREQS 8       /* frequencies per running-DFT call */

#define FDTD_OK            0
#define FDTD_EINVAL       -1   /* bad argument (size, alignment, dtype, null pointer) */
#define FDTD_ECUDA        -2   /* a CUDA runtime call or kernel launch failed */
#define FDTD_EUNSUPPORTED -3   /* valid request this build has no kernel for */

/* problem feature flags */
#define FDTD_TFSF   1          /* incident line + total-field/scattered-field corrections (programs 3_3, 3_4) */
#define FDTD_LOSSY  2          /* conduction-current integral iz / nbz (program 3_4) */
#define FDTD_ABC    4          /* 1D two-step-delay absorbing boundaries (programs 1_2 ..) */
#define FDTD_FLUX   8          /* 1D flux form Dx/Ex/Ix (programs 2_1, 2_2) */
#define FDTD_DEBYE 16          /* 1D Debye medium Sx (program 2_3) */
#define FDTD_LAZY_EZ 32        /* 2D advance: do not store ez at all (caller finishes with a non-lazy advance
                                  or fdtd2d_efield); ignored with FDTD_LOSSY.  Default: the last pass stores ez. */
#define FDTD_GHOST_DECAY 64    /* 2D advance on a slab: rows beyond the stored ones are read as zero instead of being
                                  required -- every step then invalidates one more stored row from each open end
                                  (communication-avoiding runs: k steps with k ghost rows and no exchange) */

#define FDTD_INCIDENT_READY 128 /* 2D advance with FDTD_TFSF: ezi_hist / hxi_hist already hold the incident-line history of this
                                  call's steps (fdtd2d_incident_line) -- the incident line is NOT advanced again.  For hosts that
                                  run one pass over several row blocks (streamed runs); one pass per call. */

/* structs of device pointers, in the reference's declaration order
 * (fd2d/cuda/test_3_4.cu:20-36, fd1d/cuda/test_2_3.cu:17-27) */
typedef struct { const void *fx1, *fx2, *fx3, *fy1, *fy2, *fy3, *gx2, *gx3, *gy2, *gy3; } fdtd_pmlayer;
typedef struct { const void *naz, *nbz; } fdtd_medium2d;             /* nbz may be NULL (lossless) */
typedef struct { const void *nax, *nbx, *ncx, *ndx; } fdtd_medium1d; /* ncx/ndx may be NULL */
typedef struct { void *r_pt, *i_pt, *r_in, *i_in; } fdtd_ftrans;

/* one source sample: target[index] = value (hard) or target[index] += value (soft, float64 add).
 * target == NULL: no source. */
typedef struct { void *target; long long index; int hard; double value; } fdtd_source;

/* ------------------------------------------------------------------------------------------ misc */
const char *fdtd_last_error(void);
int fdtd_version(void);
/* device facts used by the host for sizing: SM count, free / total bytes of the current device */
int fdtd_device_info(int *sm_count, size_t *free_bytes, size_t *total_bytes);
/* optional allocator / copies for hosts that do not bring their own device memory */
int fdtd_malloc(void **dptr, size_t bytes);
int fdtd_free(void *dptr);
int fdtd_memset0(void *dptr, size_t bytes, void *stream);
int fdtd_upload(void *dptr, const void *hptr, size_t bytes, void *stream);
int fdtd_download(void *hptr, const void *dptr, size_t bytes, void *stream);
int fdtd_stream_sync(void *stream);
/* CUDA IPC for the fused halo exchange (one process per GPU): export a 64-byte handle of an allocation made with
 * fdtd_malloc, open it in the neighbour's process (with ITS device current) and get a pointer its kernels can store
 * through; close before the owner frees */
int fdtd_ipc_export(const void *dptr, void *handle64);
int fdtd_ipc_open(const void *handle64, void **mapped);
int fdtd_ipc_close(void *mapped);
/* let kernels of the current device dereference memory of `peer_device` */
int fdtd_enable_peer_access(int peer_device);

/* ------------------------------------------------------------- 1D: reference-named step functions */
/* ex[1:nx] = ca*ex + cb*(hy[i-1]-hy[i]); then the source.  ca == NULL means 1, cb == NULL means 0.5.
 * replaces exfield of fd1d/cuda/test_1_1.cu:22-29, test_1_3.cu:22-30, test_1_5.cu:30-37
 * (numpy: fd1d/program/fd1d_1_5.py:65-67) */
int fdtd1d_exfield(int dtype, int nx, const void *ca, const void *cb, void *ex, const void *hy,
                   const fdtd_source *src, void *stream);
/* ABC (if abc != 0) then hy[0:nx-1] += 0.5*(ex[i]-ex[i+1]).
 * replaces hyfield of fd1d/cuda/test_1_2.cu:32-40 (numpy: fd1d/program/fd1d_2_1.py:56-61) */
int fdtd1d_hyfield(int dtype, int nx, void *ex, void *hy, void *bc, int abc, void *stream);
/* dx[1:nx] += 0.5*(hy[i-1]-hy[i]); then the source.  replaces dxfield of fd1d/cuda/test_2_3.cu:51-58 */
int fdtd1d_dxfield(int dtype, int nx, void *dx, const void *hy, const fdtd_source *src, void *stream);
/* ex = nax*(dx-ix[-ncx*sx]); ix += nbx*ex; [sx = ncx*sx + ndx*ex]   (sx == NULL: no Debye term).
 * replaces exfield of fd1d/cuda/test_2_1.cu:39-47 and test_2_3.cu:61-69 */
int fdtd1d_exfield_flux(int dtype, int nx, const fdtd_medium1d *md, const void *dx, void *ix, void *sx,
                        void *ex, void *stream);

/* running DFT: r_pt[n,:] += cos_n*ex[:], i_pt[n,:] -= sin_n*ex[:], and r_in/i_in from ex[sample_index] (10 in the
 * reference).  cosv/sinv: HOST float64 arrays of nf phase factors cos/sin(2*pi*f_n*dt*t), evaluated by the caller
 * with the reference's expression; products and sums are formed in float64 and rounded into the array type (the
 * numpy semantics for float32 arrays).  replaces fourier of fd1d/cuda/test_2_3.cu:35-48 (numpy fd1d_2_2.py:65-71) */
int fdtd1d_fourier(int dtype, int nf, int nx, const double *cosv, const double *sinv, const void *ex,
                   int sample_index, const fdtd_ftrans *ft, void *stream);

/* --------------------------------------------------------------------- 1D: fused time-blocked path */
typedef struct {
    int dtype, nx, flags;              /* FDTD_ABC | FDTD_FLUX | FDTD_DEBYE */
    const void *ca, *cb;               /* FDTD form (NULL = 1 / 0.5) */
    fdtd_medium1d md;                  /* flux form */
    void *state[2][5];                 /* two ping-pong sets of ex, hy, dx, ix, sx (unused: NULL) */
    void *bc[2];                       /* two ping-pong copies of bc[4] */
    int src_field;                     /* 0: ex, 1: dx */
    int src_index, src_hard;           /* src_index < 0: no source */
    /* Running DFT carried through the passes (programs 2_2 / 2_3 `fourier`, fd1d/program/fd1d_2_2.py:65-71): nf <= 3
     * frequencies (0 = off); ft.r_pt / ft.i_pt are nf x nx accumulators, ft.r_in / ft.i_in (nf each, may be NULL)
     * accumulate ex[dft_sample] (10 in the reference); dft_cos / dft_sin are HOST float64 tables [step][nf] of
     * cos/sin(2*pi*f*dt*t) for the steps of this call.  Every step samples Ex after the E update and before the ABC,
     * with the float64-then-round arithmetic of fdtd1d_fourier; passes are limited to 16 (float64: 8) steps. */
    int nf, dft_sample;
    fdtd_ftrans ft;
    const double *dft_cos, *dft_sin;
} fdtd1d_problem;
/* advance nsteps steps starting from state set `cur`; src[k] is the float64 waveform sample of the k-th of
 * these steps (HOST pointer, may be NULL without a source).  *cur_out = set holding the result.
 * tblock (1..64): most steps per kernel pass; passes deeper than the kernel's register budget allows (32 steps in
 * float32, 16 in float64) are split.  All arrays 16-byte aligned.  Replaces the time loop of
 * fd1d/program/fd1d_1_5.py:63-72 / fd1d_2_3.py:132-142 (CUDA: fd1d/cuda/test_1_5.cu main loop). */
int fdtd1d_advance(const fdtd1d_problem *p, int cur, int nsteps, const double *src, int tblock,
                   void *stream, int *cur_out);

/* ------------------------------------------------------------- 2D: reference-named step functions */
/* replaces ezinct of fd2d/cuda/test_3_4.cu:63-72 (numpy fd2d/program/fd2d_3_3.py:60-65) */
int fdtd2d_ezinct(int dtype, int ny, void *ezi, const void *hxi, void *bc, void *stream);
/* replaces dfield of fd2d/cuda/test_3_4.cu:75-86, test_3_2.cu:34-48 (numpy fd2d_3_3.py:68-72);
 * the embedded source assignment (ezi[3] = pulse / dz[src] = pulse) is `src` */
int fdtd2d_dfield(int dtype, int nx, int ny, const fdtd_pmlayer *pml, void *dz, const void *hx,
                  const void *hy, const fdtd_source *src, void *stream);
/* replaces inctdz of fd2d/cuda/test_3_4.cu:89-96 (numpy fd2d_3_3.py:75-78) */
int fdtd2d_inctdz(int dtype, int nx, int ny, int npml, const void *hxi, void *dz, void *stream);
/* replaces efield of fd2d/cuda/test_3_4.cu:99-109 / test_3_3.cu:69-78; iz == NULL: ez = naz*dz */
int fdtd2d_efield(int dtype, int nx, int ny, const fdtd_medium2d *md, const void *dz, void *iz, void *ez,
                  void *stream);
/* running DFT of Ez (r_pt/i_pt: nf x nx x ny) and of the source sample ezi[sample_index] (6 in the reference);
 * phase factors as for fdtd1d_fourier.  replaces fourier of fd2d/cuda/test_3_4.cu:45-60
 * (numba fd2d/python/fd2d_3_4.py:89-99).  ezi == NULL: skip r_in / i_in. */
int fdtd2d_fourier(int dtype, int nf, int nx, int ny, const double *cosv, const double *sinv, const void *ezi,
                   int sample_index, const void *ez, const fdtd_ftrans *ft, void *stream);
/* replaces hxinct of fd2d/cuda/test_3_4.cu:112-118 */
int fdtd2d_hxinct(int dtype, int ny, const void *ezi, void *hxi, void *stream);
/* replaces hfield of fd2d/cuda/test_3_4.cu:121-133 (numpy fd2d_3_3.py:91-98) */
int fdtd2d_hfield(int dtype, int nx, int ny, const fdtd_pmlayer *pml, const void *ez, void *ihx, void *ihy,
                  void *hx, void *hy, void *stream);
/* replace incthx / incthy of fd2d/cuda/test_3_4.cu:136-153 (numpy fd2d_3_3.py:101-110) */
int fdtd2d_incthx(int dtype, int nx, int ny, int npml, const void *ezi, void *hx, void *stream);
int fdtd2d_incthy(int dtype, int nx, int ny, int npml, const void *ezi, void *hy, void *stream);

/* setup on the device: naz / nbz of the lossy dielectric cylinder of program 3_4 for GLOBAL rows [row_lo, row_hi)
 * (arrays of (row_hi-row_lo) x ny), float64 evaluation of the reference's Python statements
 * (fd2d/python/fd2d_3_4.py:173-194), bit-identical to them */
int fdtd2d_dielectric_cylinder(int dtype, int nx, int ny, int npml, int rgrid, double dt, double epsr, double sigma,
                               int row_lo, int row_hi, void *naz, void *nbz, void *stream);

/* setup on the device: the ten PML vectors of pmlparam (fd2d/program/fd2d_3_3.py:113-122; defaults :147-158) written
 * into the arrays `pml` points to (x-vectors: nx entries, y-vectors: ny), float64 evaluation of the reference's Python
 * statements rounded on store.  npml = 0: the identity set (free space).
 * `cubes`: NULL, or 2*npml doubles ON THE DEVICE holding the host's ((npml-n)/npml)**3, n = 0..npml-1, followed by
 * ((npml-n-0.5)/npml)**3 -- Python's ** is the host libm's pow(), which is within one ulp but not correctly rounded, so
 * float64 vectors are bit-identical to the reference's only with the host's own cubes.  With NULL the kernel cubes in
 * double-double (correctly rounded): float32 vectors still come out bit-identical, float64 ones to one ulp. */
int fdtd2d_pmlparam(int dtype, int nx, int ny, int npml, const double *cubes, const fdtd_pmlayer *pml, void *stream);

/* --------------------------------------------------------------------- 2D: fused time-blocked path */
enum { FDTD2D_DZ = 0, FDTD2D_EZ, FDTD2D_HX, FDTD2D_HY, FDTD2D_IHX, FDTD2D_IHY, FDTD2D_IZ, FDTD2D_NFIELDS };

typedef struct {
    int dtype;
    int nx, ny;                 /* GLOBAL grid size */
    int row_lo, row_hi;         /* global rows this device owns and must produce: [row_lo, row_hi) */
    int row_base, rows_alloc;   /* global row index of array row 0, and rows stored per array (owned + ghosts) */
    int npml, flags;            /* FDTD_TFSF | FDTD_LOSSY */
    fdtd_pmlayer pml;           /* x-vectors: GLOBAL length nx; y-vectors: length ny */
    fdtd_medium2d md;           /* local layout (rows_alloc x ny) */
    void *state[2][FDTD2D_NFIELDS];  /* two ping-pong sets, local layout; IZ may be NULL without FDTD_LOSSY */
    void *ezi, *hxi, *bc;       /* incident line, length ny / ny / 4 (FDTD_TFSF) */
    void *ezi_hist, *hxi_hist;  /* scratch: tblock*ny and tblock*2 elements (FDTD_TFSF) */
    int src_i, src_j, src_hard; /* point source on dz at GLOBAL (src_i, src_j); src_i < 0: none.
                                   With FDTD_TFSF the table drives ezi[3] instead (hard). */
    /* Optional promise that unlocks the interior fast kernel: for GLOBAL rows i in [ident_row_lo, ident_row_hi)
     * the x-vectors, and for columns j in [ident_col_lo, ident_col_hi) the y-vectors, hold exactly the
     * identity set (f?1 = 0, f?2 = f?3 = g?2 = g?3 = 1) -- true for [npml, N-1-npml) after pmlparam
     * (fd2d/program/fd2d_3_3.py:113-122).  Multiplications by 1 are then skipped (exact).  All zero: no promise.
     * fdtd2d_check_identity() verifies a promise on the device. */
    int ident_row_lo, ident_row_hi, ident_col_lo, ident_col_hi;
    /* Running DFT fused into the passes (program 3_4's `fourier`, fd2d/python/fd2d_3_4.py:89-99): nf <= 3
     * frequencies (0 = off); ft.r_pt / ft.i_pt are nf x rows_alloc x ny accumulators, ft.r_in / ft.i_in (nf each,
     * FDTD_TFSF only, may be NULL) accumulate the source sample ezi[6]; dft_cos / dft_sin are HOST float64 tables
     * [step][nf] of cos/sin(2*pi*f*dt*t) for the steps of this call.  Same float64-then-round arithmetic and the same
     * per-step order as fdtd2d_fourier after every step; passes are limited to 4 steps and 2-wide vectors. */
    int nf;
    fdtd_ftrans ft;
    const double *dft_cos, *dft_sin;
    /* Halo exchange fused into the pass (multi-GPU slabs, one process per GPU; all zero = off).  The last pass of
     * the call ALSO stores its first / last `halo` owned rows (dz, hx, hy, ihx, ihy[, iz]) into the neighbours'
     * ghost rows through peer-mapped pointers (cudaIpcOpenMemHandle), row by row while it computes, and the last
     * warp of the pass publishes `epoch` in the neighbours' sync words; the first pass of the next call spins on
     * its own sync words until both neighbours have published epoch-1.  peer_up / peer_dn: the neighbours' two
     * array sets (entries NULL where there is no neighbour); *_base: global row of their array row 0; sync_*: EIGHT
     * zero-initialised 64-bit words per rank {from_up, from_down, counter, counter, error, reserved x3}; epoch: 1, 2,
     * ... per call, the same on every rank.  nsteps <= halo.  The wait is bounded (20 s by default, fdtd2d_tune2): a
     * pass whose neighbour never arrives gives up, raises the error word and ends; fdtd2d_halo_status() reports it. */
    int halo;
    void *peer_up[2][FDTD2D_NFIELDS], *peer_dn[2][FDTD2D_NFIELDS];
    int peer_up_base, peer_dn_base;
    void *sync_local, *sync_up, *sync_dn;
    unsigned long long epoch;
    /* FDTD_LOSSY problems whose loss is local (the reference's dielectric cylinder in free space, fd2d/python/
     * fd2d_3_4.py:173-194): a promise that OUTSIDE global rows [lossy_row_lo, lossy_row_hi) x columns [lossy_col_lo,
     * lossy_col_hi) nbz is 0 and iz is +0 in both state sets.  There `ez = naz*(dz-iz); iz += nbz*ez` has the bits of
     * `ez = naz*dz` and leaves iz alone, so interior warps that stay outside the box run the lossless kernel (no iz /
     * nbz traffic).  All four zero (or an empty box): no promise.  fdtd2d_check_lossless_outside() verifies it. */
    int lossy_row_lo, lossy_row_hi, lossy_col_lo, lossy_col_hi;
} fdtd2d_problem;

/* Advance nsteps full time steps (reference order: ezinct, dfield+source, inctdz, efield, hxinct, hfield,
 * incthx, incthy -- fd2d/python/fd2d_3_4.py:268-277) from state set `cur`, at most tblock (1..12) steps per kernel
 * pass; tblock = 0 lets the library choose depth, vector width and chunking by grid size and step count
 * (fdtd2d_plan returns the choice).
 * Rows [row_lo-g, row_hi+g) with g = nsteps must be present and current in set `cur` (clipped to the
 * grid); single device: row_lo=row_base=0, row_hi=rows_alloc=nx and any nsteps is allowed.
 * src: HOST float64 table, one sample per step.  *cur_out = set holding the result. */
int fdtd2d_advance(const fdtd2d_problem *p, int cur, int nsteps, const double *src, int tblock,
                   void *stream, int *cur_out);
/* number of coefficient entries that violate the ident_* promise of `p` (0 = promise holds); synchronises */
int fdtd2d_check_identity(const fdtd2d_problem *p, long long *violations);
/* *violations = stored cells outside the lossless-outside box of `p` with nbz != 0 or iz != +0 (either set) */
int fdtd2d_check_lossless_outside(const fdtd2d_problem *p, long long *violations);
/* resolve (module-load) every kernel instantiation a problem of this dtype / width can launch, so that no
 * time step pays CUDA's lazy loading.  lossy: bit 0 = FDTD_LOSSY kernels, bit 1 = also the fused-DFT kernels */
int fdtd2d_preload(int dtype, int ny, int lossy);
/* largest supported tblock for a dtype / ny (0 if unsupported) */
int fdtd2d_max_tblock(int dtype, int ny);
/* The launch plan fdtd2d_advance(p, cur, nsteps, src, tblock, ...) will use, without launching anything: the pass
 * depths in order (up to `cap` of them are written to depths[], which may be NULL), the vector width and the rows per
 * chunk of the first pass.  Only dtype, nx, ny, row_lo, row_hi, flags and nf of `p` are read.  Returns the number of
 * passes (>= 0) or a negative FDTD_E* code.  Hosts that must know the depth in advance (ghost-band width of a slab,
 * pass levels of a streamed run) ask here instead of re-deriving it. */
int fdtd2d_plan(const fdtd2d_problem *p, int nsteps, int tblock, int *depths, int cap, int *vector_width,
                int *chunk_rows);
/* Advance the TFSF incident line of `p` (ezi, hxi, bc; ezinct ... hxinct with the hard source ezi[3] = src[k],
 * fd2d/program/fd2d_3_3.py:60-65,72,86-88) by nsteps <= 12 steps and record what a pass over those steps needs: ezi after
 * ezinct + source of every step (ezi_hist: nsteps x ny) and hxi[npml-2], hxi[ny-npml] before hxinct (hxi_hist: nsteps x 2).
 * fdtd2d_advance does this itself once per pass; hosts that cut one pass into row blocks call it once per pass and hand
 * the history to every block's call with FDTD_INCIDENT_READY. */
int fdtd2d_incident_line(const fdtd2d_problem *p, int nsteps, const double *src, void *ezi_hist, void *hxi_hist,
                         void *stream);
/* Fused halo exchange: *word = 0 while every wait of every pass so far was answered; otherwise epoch*4 + side (1 = the
 * upper, 2 = the lower neighbour never arrived within the bound) and fdtd_last_error() says so.  Synchronises. */
int fdtd2d_halo_status(const fdtd2d_problem *p, unsigned long long *word);
/* Test / tuning hook, process-wide, all zero = defaults: force_v (1, 2, 4) and chunk_rows override the launch plan,
 * warps_per_cta (1..8), ring_depth == 1 runs the edge and interior kernels in stream order instead of forking the
 * edge kernel onto a side stream, force_careful bit 0 sends every warp through the careful (edge) kernel, bit 1
 * ignores the lossless-outside promise.  Results are bit-identical under every setting (that is what the tests use
 * it for). */
int fdtd2d_tune(int force_v, int chunk_rows, int warps_per_cta, int ring_depth, int force_careful);
/* More process-wide knobs, by key.  FDTD_TUNE_DEEP: 1 (default) = passes of depth 8 and 12 use the deep kernels where
 * they apply (float, 4-wide vectors, no fused DFT), 0 = never (depth <= 8, register-pipeline kernels only), 2 =
 * additionally run the shared-memory-resident careful (edge) kernel at every depth.  FDTD_TUNE_HALO_WAIT_MS: bound of
 * the fused halo exchange's wait for a neighbour in milliseconds (0 = the default of 20 s).  FDTD_TUNE_VARIANT: interior
 * kernel of the deep passes -- 0 (default) = the warp-chain kernel (TMA-fed pipeline of warps; shipped shape), 1..3 =
 * the shared-memory-accumulator kernels, >= 10 = other warp-chain shapes.  Results are bit-identical under every
 * setting. */
#define FDTD_TUNE_DEEP 0
#define FDTD_TUNE_HALO_WAIT_MS 1
#define FDTD_TUNE_VARIANT 2
#define FDTD_TUNE_COL_FAST 4         /* bit 0: PML-column strips, bit 1: PML-row chunks of a deep pass through the warp-chain kernel's column / row variant (default 3); 0 = edge kernel */
#define FDTD_TUNE_EDGE_CHUNKS 3      /* 1 (default) = short first / last row chunk around the rows that need the edge kernel; 0 = uniform */
int fdtd2d_tune2(int key, long long value);

#ifdef __cplusplus
}
#endif
#endif /* FDTD_B200_H */
