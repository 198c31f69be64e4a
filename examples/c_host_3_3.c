/* C host for program 3_3 (TFSF plane wave + PML) on libfdtd_b200 -- the reference's fd2d/cuda/test_3_3.cu main()
 * with its <<<>>> launches replaced by the same-named C-ABI calls (unfused loop), then the same run through
 * fdtd2d_advance (fused, time-blocked), and a byte comparison of the two Ez fields.  Prints ez[2][0:50] like the
 * reference benchmark.   Build:  gcc examples/c_host_3_3.c -Iinclude -Lsimulation_b200/csrc -lfdtd_b200 -lm
 * Run:    LD_LIBRARY_PATH=simulation_b200/csrc ./a.out [nx ny ns npml] */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fdtd_b200.h"

#define CHECK(call)                                                              \
    do {                                                                         \
        if ((call) != FDTD_OK) {                                                 \
            fprintf(stderr, "%s failed: %s\n", #call, fdtd_last_error());        \
            return 1;                                                            \
        }                                                                        \
    } while (0)

static float *dev_zeros(size_t n) {
    void *p = NULL;
    if (fdtd_malloc(&p, n * sizeof(float)) != FDTD_OK || fdtd_memset0(p, n * sizeof(float), NULL) != FDTD_OK) exit(2);
    return (float *)p;
}

static float *dev_from(const float *h, size_t n) {
    float *d = dev_zeros(n);
    if (fdtd_upload(d, h, n * sizeof(float), NULL) != FDTD_OK || fdtd_stream_sync(NULL) != FDTD_OK) exit(2);
    return d;
}

/* the reference's pmlparam (fd2d/program/fd2d_3_3.py:113-122), evaluated in double and rounded on store */
static void pmlparam(int nx, int ny, int npml, float **v /* fx1 fx2 fx3 fy1 fy2 fy3 gx2 gx3 gy2 gy3 */) {
    const int len[10] = {nx, nx, nx, ny, ny, ny, nx, nx, ny, ny};
    for (int k = 0; k < 10; k++) {
        v[k] = (float *)malloc(len[k] * sizeof(float));
        for (int i = 0; i < len[k]; i++) v[k][i] = (k == 0 || k == 3) ? 0.0f : 1.0f;
    }
    for (int n = 0; n < npml; n++) {
        double xm = 0.33 * pow((double)(npml - n) / npml, 3), xn = 0.33 * pow((npml - n - 0.5) / npml, 3);
        for (int ax = 0; ax < 2; ax++) {
            int N = ax ? ny : nx;
            float *f1 = v[ax ? 3 : 0], *f2 = v[ax ? 4 : 1], *f3 = v[ax ? 5 : 2], *g2 = v[ax ? 8 : 6], *g3 = v[ax ? 9 : 7];
            f1[n] = f1[N - 2 - n] = (float)xn;
            f2[n] = f2[N - 2 - n] = (float)(1 / (1 + xn));
            f3[n] = f3[N - 2 - n] = (float)((1 - xn) / (1 + xn));
            g2[n] = g2[N - 1 - n] = (float)(1 / (1 + xm));
            g3[n] = g3[N - 1 - n] = (float)((1 - xm) / (1 + xm));
        }
    }
}

int main(int argc, char **argv) {
    int nx = argc > 1 ? atoi(argv[1]) : 1024, ny = argc > 2 ? atoi(argv[2]) : 1024;
    int ns = argc > 3 ? atoi(argv[3]) : 500, npml = argc > 4 ? atoi(argv[4]) : 80;
    size_t n2 = (size_t)nx * ny;
    float *hv[10], *dv[10];
    pmlparam(nx, ny, npml, hv);
    const int len[10] = {nx, nx, nx, ny, ny, ny, nx, nx, ny, ny};
    for (int k = 0; k < 10; k++) dv[k] = dev_from(hv[k], len[k]);
    fdtd_pmlayer pml = {dv[0], dv[1], dv[2], dv[3], dv[4], dv[5], dv[6], dv[7], dv[8], dv[9]};
    float *ones = (float *)malloc(n2 * sizeof(float));
    for (size_t i = 0; i < n2; i++) ones[i] = 1.0f;
    float *naz = dev_from(ones, n2);
    fdtd_medium2d md = {naz, NULL};
    double *src = (double *)malloc(ns * sizeof(double));
    for (int t = 1; t <= ns; t++) src[t - 1] = exp(-0.5 * pow((t - 20) / 8.0, 2));   /* gaussian(t, 20, 8.0) */

    /* ---- (a) the reference loop, one call per reference function */
    float *ezi = dev_zeros(ny), *hxi = dev_zeros(ny), *bc = dev_zeros(4);
    float *dz = dev_zeros(n2), *ez = dev_zeros(n2), *hx = dev_zeros(n2), *hy = dev_zeros(n2), *ihx = dev_zeros(n2),
          *ihy = dev_zeros(n2);
    for (int t = 1; t <= ns; t++) {
        fdtd_source s = {ezi, 3, 1, src[t - 1]};
        CHECK(fdtd2d_ezinct(FDTD_F32, ny, ezi, hxi, bc, NULL));
        CHECK(fdtd2d_dfield(FDTD_F32, nx, ny, &pml, dz, hx, hy, &s, NULL));
        CHECK(fdtd2d_inctdz(FDTD_F32, nx, ny, npml, hxi, dz, NULL));
        CHECK(fdtd2d_efield(FDTD_F32, nx, ny, &md, dz, NULL, ez, NULL));
        CHECK(fdtd2d_hxinct(FDTD_F32, ny, ezi, hxi, NULL));
        CHECK(fdtd2d_hfield(FDTD_F32, nx, ny, &pml, ez, ihx, ihy, hx, hy, NULL));
        CHECK(fdtd2d_incthx(FDTD_F32, nx, ny, npml, ezi, hx, NULL));
        CHECK(fdtd2d_incthy(FDTD_F32, nx, ny, npml, ezi, hy, NULL));
    }
    float *ez_a = (float *)malloc(n2 * sizeof(float));
    CHECK(fdtd_download(ez_a, ez, n2 * sizeof(float), NULL));
    CHECK(fdtd_stream_sync(NULL));

    /* ---- (b) the fused path: the whole loop in one call */
    fdtd2d_problem p;
    memset(&p, 0, sizeof(p));
    p.dtype = FDTD_F32; p.nx = nx; p.ny = ny; p.row_lo = 0; p.row_hi = nx; p.row_base = 0; p.rows_alloc = nx;
    p.npml = npml; p.flags = FDTD_TFSF; p.pml = pml; p.md = md;
    for (int s = 0; s < 2; s++)
        for (int f = 0; f < FDTD2D_IZ; f++) p.state[s][f] = dev_zeros(n2);
    p.ezi = dev_zeros(ny); p.hxi = dev_zeros(ny); p.bc = dev_zeros(4);
    p.ezi_hist = dev_zeros((size_t)8 * ny); p.hxi_hist = dev_zeros(16);
    p.src_i = -1;
    p.ident_row_lo = npml; p.ident_row_hi = nx - 1 - npml; p.ident_col_lo = npml; p.ident_col_hi = ny - 1 - npml;
    long long bad = -1;
    CHECK(fdtd2d_check_identity(&p, &bad));
    if (bad != 0) { fprintf(stderr, "identity promise violated (%lld)\n", bad); return 1; }
    int cur = 0;
    CHECK(fdtd2d_advance(&p, 0, ns, src, 0, NULL, &cur));
    float *ez_b = (float *)malloc(n2 * sizeof(float));
    CHECK(fdtd_download(ez_b, p.state[cur][FDTD2D_EZ], n2 * sizeof(float), NULL));
    CHECK(fdtd_stream_sync(NULL));

    int same = memcmp(ez_a, ez_b, n2 * sizeof(float)) == 0;
    float peak = 0;
    for (size_t i = 0; i < n2; i++) if (fabsf(ez_a[i]) > peak) peak = fabsf(ez_a[i]);
    printf("program 3_3 %dx%d ns=%d npml=%d: peak |ez| = %.6f, fused == unfused: %s\n", nx, ny, ns, npml, peak,
           same ? "identical bytes" : "DIFFERENT");
    for (int j = 0; j < 50 && j < ny; j++) printf("%.8e%s", ez_b[2 * (size_t)ny + j], (j % 6 == 5) ? "\n" : " ");
    printf("\n");
    return same ? 0 : 1;
}
