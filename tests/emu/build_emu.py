"""TEST INFRASTRUCTURE ONLY.  Compiles a kernel source file of simulation_b200/csrc for the HOST against the stand-in
tests/emu/cuda_runtime.h: every ``kernel<<<grid, block, smem, stream>>>(args)`` is rewritten into
``emu::launch(grid, block, [&]{ kernel(args); })`` and the result is built with g++ into a shared object under
tests/emu/_build/.  Used by tests/test_emu_*.py to run the kernels' own source (indexing, masks, ownership, rounding
order) on a machine without a GPU.  Nothing under simulation_b200/ may import this."""
from __future__ import annotations

import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "simulation_b200", "csrc")

STUBS = r'''
// ---- what the other translation units of the library provide
namespace fdtd {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap); }
int cuda_fail(cudaError_t e, const char *what, const char *, int) { set_error("emulated CUDA error %d: %s", (int)e, what); return FDTD_ECUDA; }
int sm_count() { return 148; }
int launch_fourier(int, int, size_t, const double *, const double *, const void *, const void *, const fdtd_ftrans *, cudaStream_t) {
    set_error("launch_fourier is not part of the emulated build"); return FDTD_EUNSUPPORTED; }
}
extern "C" const char *fdtd_last_error(void) { return fdtd::g_err; }
extern "C" long long emu_launches(void) { return emu::launches; }
'''


def _match(text, i, open_ch, close_ch):
    """index just past the bracket that closes text[i] (== open_ch)"""
    depth = 0
    for k in range(i, len(text)):
        if text[k] == open_ch:
            depth += 1
        elif text[k] == close_ch:
            depth -= 1
            if depth == 0:
                return k + 1
    raise ValueError("unbalanced")


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def rewrite_launches(src: str) -> str:
    out, pos = "", 0
    while True:
        i = src.find("<<<", pos)
        if i < 0:
            return out + src[pos:]
        # kernel expression: identifier, optionally followed by <template arguments>
        k = i
        if src[k - 1] == ">":
            depth = 0
            while True:
                k -= 1
                if src[k] == ">":
                    depth += 1
                elif src[k] == "<":
                    depth -= 1
                    if depth == 0:
                        break
        while k > 0 and (src[k - 1].isalnum() or src[k - 1] in "_:"):
            k -= 1
        kernel = src[k:i]
        j = src.index(">>>", i)
        cfg = _split_top(src[i + 3:j])
        a0 = src.index("(", j)
        a1 = _match(src, a0, "(", ")")
        out += src[pos:k] + f"emu::launch(dim3({cfg[0]}), dim3({cfg[1]}), [&] {{ {kernel}{src[a0:a1]}; }})"
        pos = a1


def build(source_name: str) -> str:
    """-> path of the emulated shared object for simulation_b200/csrc/<source_name>"""
    src = open(os.path.join(CSRC, source_name)).read()
    hdr = open(os.path.join(HERE, "cuda_runtime.h")).read() + open(os.path.join(CSRC, "common.cuh")).read()
    tag = hashlib.sha256((src + hdr + STUBS).encode()).hexdigest()[:16]
    bdir = os.path.join(HERE, "_build")
    os.makedirs(bdir, exist_ok=True)
    so = os.path.join(bdir, f"emu_{os.path.splitext(source_name)[0]}_{tag}.so")
    if os.path.exists(so):
        return so
    cpp = os.path.join(bdir, f"emu_{os.path.splitext(source_name)[0]}.cpp")
    with open(cpp, "w") as f:
        f.write(f'#line 1 "{source_name}"\n' + rewrite_launches(src) + STUBS)
    cmd = ["g++", "-std=c++20", "-O1", "-g", "-ffp-contract=off", "-fno-strict-aliasing", "-fPIC", "-shared", "-pthread", "-w",
           "-I", HERE, "-I", CSRC, "-x", "c++", cpp, "-o", so + ".tmp"]
    subprocess.run(cmd, check=True)
    os.replace(so + ".tmp", so)
    return so


if __name__ == "__main__":
    import sys
    print(build(sys.argv[1] if len(sys.argv) > 1 else "fd1d.cu"))
