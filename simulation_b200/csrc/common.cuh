// Shared helpers for libfdtd_b200 (sm_100a).  Build with -fmad=false: every kernel spells out the
// reference's left-to-right evaluation order and relies on the compiler NOT contracting a*b+c.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/fdtd_b200.h"

namespace fdtd {

void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define FDTD_CUDA(call)                                                                          \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess) return fdtd::cuda_fail(e__, #call, __FILE__, __LINE__);          \
    } while (0)

#define FDTD_LAUNCH_CHECK(name)                                                                  \
    do {                                                                                         \
        cudaError_t e__ = cudaGetLastError();                                                    \
        if (e__ != cudaSuccess) return fdtd::cuda_fail(e__, name, __FILE__, __LINE__);           \
    } while (0)

#define FDTD_REQUIRE(cond, ...)                                                                  \
    do {                                                                                         \
        if (!(cond)) { fdtd::set_error(__VA_ARGS__); return FDTD_EINVAL; }                       \
    } while (0)

inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }
inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int sm_count();   // cached SM count of the current device
int launch_fourier(int dtype, int nf, size_t ncells, const double *cosv, const double *sinv, const void *field,
                   const void *sample, const fdtd_ftrans *ft, cudaStream_t st);

// Source injection with the numpy semantics: hard = rounded store, soft = float64 add then one rounding.
template <typename real>
__device__ __forceinline__ real inject(real old, double value, int hard) {
    return hard ? static_cast<real>(value) : static_cast<real>(static_cast<double>(old) + value);
}

}  // namespace fdtd
