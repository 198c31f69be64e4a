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

struct ChainMaps { CUtensorMap m[NARR]; };   // in_dz in_hx in_hy in_ihx in_ihy naz: (rows, ny) float, box R x 128

template <int G_, int K_, int R_, int NSTAGE_, int QD_, int GROUPS_>
struct ChainShape {
    static constexpr int G = G_, K = K_, R = R_, NSTAGE = NSTAGE_, QD = QD_, GROUPS = GROUPS_;
    static constexpr int T = G * K, NS = K + 1;
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
};

// ---- mbarrier and TMA in PTX (sm_90+)
__device__ __forceinline__ unsigned smem_u32(const void *p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(const unsigned bar, const unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(const unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(const unsigned bar, const unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(const unsigned bar, const unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "CHAIN_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra CHAIN_DONE;\n"
        "bra CHAIN_WAIT;\n"
        "CHAIN_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// one box (R rows x 128 columns of one array) global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_2d(const unsigned dst, const CUtensorMap *map, const unsigned bar, const int col,
                                            const int row) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(col), "r"(row) : "memory");
}

__device__ __forceinline__ void sts4(void *dst, const float (&d)[CV]) {
    *reinterpret_cast<float4 *>(dst) = make_float4(d[0], d[1], d[2], d[3]);
}

template <typename Shape>
__device__ __forceinline__ void chain_body(const MarchParams<float> &p, const ChainMaps &maps, const int strip, const int i0,
                                           const int i1, const int lane, const int wg, unsigned char *const gsm) {
    constexpr int G = Shape::G, K = Shape::K, R = Shape::R, NSTAGE = Shape::NSTAGE, QD = Shape::QD;
    constexpr int T = Shape::T, NS = Shape::NS, HALO = Shape::HALO, W = Shape::W, USE = Shape::USE;
    const bool first = wg == 0, last = wg == G - 1;

    const int c0 = strip * USE - HALO;               // first column of the strip (halo included)
    const int jb = c0 + lane * CV;                   // first column of this lane
    const bool col_store = (lane * CV >= HALO) && (lane * CV + CV <= W - HALO);
    const int r_begin = i0 - T, r_end = i1 + T;      // rows fed to stage 0: [r_begin, r_end)
    const int n_in = (r_end - r_begin) - wg * K;     // rows this warp receives

    // shared memory of the group
    unsigned char *const stage = gsm;                                    // [NSTAGE][NARR][R][CROWB]
    unsigned char *const queue = gsm + Shape::OFF_Q;                     // [G-1][QD][NARR][CROWB]
    const unsigned bars = smem_u32(gsm + Shape::OFF_BAR);
    // barrier ids: stage full s -> s; queue q slot d: full -> NSTAGE + (q*QD + d)*2, empty -> ... + 1
    auto bar_stage = [&](const int s) { return bars + 8u * s; };
    auto bar_qfull = [&](const int q, const int d) { return bars + 8u * (NSTAGE + (q * QD + d) * 2); };
    auto bar_qempty = [&](const int q, const int d) { return bars + 8u * (NSTAGE + (q * QD + d) * 2 + 1); };

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

    // ---- first warp: the TMA side.  Box b holds the rows r_begin + b*R .. + R - 1 of the six arrays.
    const int n_box = (n_in + R - 1) / R;            // (meaningful for the first warp)
    auto issue_box = [&](const int b, const int slot) {       // one lane
        const unsigned bar = bar_stage(slot);
        mbar_arrive_expect_tx(bar, (unsigned)Shape::STAGE_B);
        const unsigned dst = smem_u32(stage + slot * Shape::STAGE_B);
        const int row = r_begin + b * R - p.row_base;
#pragma unroll
        for (int a = 0; a < NARR; ++a) tma_load_2d(dst + a * Shape::BOX_B, &maps.m[a], bar, c0, row);
    };
    if (first && lane == 0) {
#pragma unroll
        for (int b = 0; b < NSTAGE; ++b)
            if (b < n_box) issue_box(b, b);
    }

    int in_slot = 0, in_row = 0, box = 0;            // input ring position (first warp: staging slot / row in box / box id)
    unsigned in_phase = 0;
    int out_slot = 0;
    unsigned out_phase = 0;
    long long off_s = (long long)(r_begin - K - p.row_base) * p.ny + jb;   // last warp: element offset of the row stored next

    auto st2 = [&](float *dst, const float (&d)[CV]) {      // register pairs leave as two 64-bit halves (no quad assembly)
        *reinterpret_cast<float2 *>(dst) = make_float2(d[0], d[1]);
        *reinterpret_cast<float2 *>(dst + 2) = make_float2(d[2], d[3]);
    };

#pragma unroll 1
    for (int x0 = 0; x0 < n_in; x0 += NS) {
#pragma unroll
        for (int u = 0; u < NS; ++u) {
            const int x = x0 + u;                     // this warp's x-th input: global row r_begin + x
            if (x >= n_in) break;
            // ---- take the row into register set u
            {
                const unsigned char *src;
                int astride;
                if (first) {
                    if (in_row == 0) mbar_wait(bar_stage(in_slot), in_phase);
                    src = stage + in_slot * Shape::STAGE_B + in_row * CROWB + lane * CLB;
                    astride = Shape::BOX_B;
                } else {
                    mbar_wait(bar_qfull(wg - 1, in_slot), in_phase);
                    src = queue + ((wg - 1) * QD + in_slot) * Shape::QSLOT_B + lane * CLB;
                    astride = CROWB;
                }
                lds_vec<float, CV>(src + 0 * astride, S[u].dz);
                lds_vec<float, CV>(src + 1 * astride, S[u].hx);
                lds_vec<float, CV>(src + 2 * astride, S[u].hy);
                lds_vec<float, CV>(src + 3 * astride, S[u].ihx);
                lds_vec<float, CV>(src + 4 * astride, S[u].ihy);
                lds_vec<float, CV>(src + 5 * astride, S[u].naz);
            }
            // ---- this warp's K stages: stage s has row x-s arriving and holds row x-s-1
#pragma unroll
            for (int s = 0; s < K; ++s) {
                const int sa = (u - s + 2 * NS) % NS, sh = (u - s - 1 + 2 * NS) % NS;
                march_stage_pk<CV, false, false>(S[sa], S[sh], negzero, nullptr, nullptr);
            }
            // ---- give the input slot back (every value read from it has been used by now)
            __syncwarp();
            if (first) {
                if (++in_row == R || x + 1 == n_in) {
                    in_row = 0;
                    if (lane == 0 && box + NSTAGE < n_box) issue_box(box + NSTAGE, in_slot);
                    ++box;
                    if (++in_slot == NSTAGE) { in_slot = 0; in_phase ^= 1u; }
                }
            } else {
                if (lane == 0) mbar_arrive(bar_qempty(wg - 1, in_slot));
                if (++in_slot == QD) { in_slot = 0; in_phase ^= 1u; }
            }
            // ---- the row leaving the last stage: set (u+1) % NS = this warp's input x-K, now K levels later
            if (x >= K) {
                const RowSet<float, CV> &O = S[(u + 1) % NS];
                if (last) {
                    const int ro = r_begin + x - K;
                    if (ro >= i0 && ro < i1 && col_store) {
                        st2(p.out_dz + off_s, O.dz);
                        st2(p.out_hx + off_s, O.hx);
                        st2(p.out_hy + off_s, O.hy);
                        st2(p.out_ihx + off_s, O.ihx);
                        st2(p.out_ihy + off_s, O.ihy);
                        if (p.write_ez) st2(p.out_ez + off_s, O.ez);
                    }
                } else {
                    mbar_wait(bar_qempty(wg, out_slot), out_phase ^ 1u);
                    unsigned char *dst = queue + (wg * QD + out_slot) * Shape::QSLOT_B + lane * CLB;
                    sts4(dst + 0 * CROWB, O.dz);
                    sts4(dst + 1 * CROWB, O.hx);
                    sts4(dst + 2 * CROWB, O.hy);
                    sts4(dst + 3 * CROWB, O.ihx);
                    sts4(dst + 4 * CROWB, O.ihy);
                    sts4(dst + 5 * CROWB, O.naz);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_qfull(wg, out_slot));
                    if (++out_slot == QD) { out_slot = 0; out_phase ^= 1u; }
                }
            }
            off_s += p.ny;
        }
    }
}

template <typename Shape>
__global__ void __launch_bounds__(Shape::THREADS, 1)
k_march_chain(const __grid_constant__ MarchParams<float> p, const __grid_constant__ ChainMaps maps) {
    extern __shared__ __align__(1024) unsigned char chain_smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int grp = warp / Shape::G, wg = warp % Shape::G;
    unsigned char *const gsm = chain_smem + (size_t)grp * Shape::GROUP_SMEM;
    if (wg == 0 && lane == 0) {
        const unsigned bars = smem_u32(gsm + Shape::OFF_BAR);
        for (int b = 0; b < Shape::NBAR; ++b) mbar_init(bars + 8u * b, 1u);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    int strip, i0, i1;
    if (!decode_item<true>(p, blockIdx.x * Shape::GROUPS + grp, 0, CV, Shape::T, false, strip, i0, i1)) return;   // the whole group
    chain_body<Shape>(p, maps, strip, i0, i1, lane, wg, gsm);
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

template <typename Shape>
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
            FDTD_CUDA(cudaFuncSetAttribute(k_march_chain<Shape>, cudaFuncAttributeMaxDynamicSharedMemorySize, Shape::SMEM));
            done = true;
        }
    }
    const int grid = (items + Shape::GROUPS - 1) / Shape::GROUPS;
    k_march_chain<Shape><<<grid, Shape::THREADS, Shape::SMEM, st>>>(mp, maps);
    FDTD_LAUNCH_CHECK("k_march_chain");
    return FDTD_OK;
}

//                        G  K  R  NSTAGE QD GROUPS
using Chain8 = ChainShape<4, 2, 2, 3, 2, 4>;          // depth 8: sixteen warps per SM
using Chain8b = ChainShape<4, 2, 4, 2, 3, 4>;         // ... bigger boxes, deeper queues
using Chain8c = ChainShape<2, 4, 2, 3, 2, 4>;         // ... two warps of four stages (eight warps per SM)
using Chain12 = ChainShape<4, 3, 2, 3, 2, 3>;         // depth 12: twelve warps per SM
using Chain12b = ChainShape<6, 2, 2, 3, 2, 2>;        // ... six warps of two stages

}  // namespace

namespace fdtd_march {

bool chain_supported(int T, bool lossy) {
    return (T == 8 || T == 12) && !lossy && encoder() != nullptr;
}

int launch_march_chain(const MarchParams<float> &mp, int T, int shape, int items, cudaStream_t st) {
    if (mp.ny % CV != 0 || (reinterpret_cast<uintptr_t>(mp.in_dz) & 15u) != 0) {
        fdtd::set_error("warp-chain pass: ny must be a multiple of 4 and the arrays 16-byte aligned");
        return FDTD_EINVAL;
    }
    if (T == 8) {
        if (shape == 1) return launch_chain<Chain8b>(mp, items, st);
        if (shape == 2) return launch_chain<Chain8c>(mp, items, st);
        return launch_chain<Chain8>(mp, items, st);
    }
    if (T == 12) {
        if (shape == 1) return launch_chain<Chain12b>(mp, items, st);
        return launch_chain<Chain12>(mp, items, st);
    }
    fdtd::set_error("no warp-chain pass of depth %d", T);
    return FDTD_EUNSUPPORTED;
}

void preload_chain() {
    cudaFuncAttributes a;
    cudaFuncGetAttributes(&a, k_march_chain<Chain8>);
    cudaFuncGetAttributes(&a, k_march_chain<Chain12>);
}

}  // namespace fdtd_march
