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
extern "C" {
const char *fdtd_last_error(void) { return fdtd::g_err; }
long long emu_launches(void) { return emu::launches; }
// capi_misc.cu on host memory ("device pointers" are host pointers here); no peers, no IPC
int fdtd_version(void) { return 100; }
int fdtd_device_info(int *sm, size_t *f, size_t *t) { if (sm) *sm = 148; if (f) *f = 0; if (t) *t = 0; return FDTD_OK; }
int fdtd_malloc(void **p, size_t n) { *p = std::aligned_alloc(256, (n + 255) / 256 * 256); return *p ? FDTD_OK : FDTD_ECUDA; }
int fdtd_free(void *p) { std::free(p); return FDTD_OK; }
int fdtd_memset0(void *p, size_t n, void *) { std::memset(p, 0, n); return FDTD_OK; }
int fdtd_upload(void *d, const void *h, size_t n, void *) { std::memcpy(d, h, n); return FDTD_OK; }
int fdtd_download(void *h, const void *d, size_t n, void *) { std::memcpy(h, d, n); return FDTD_OK; }
int fdtd_stream_sync(void *) { return FDTD_OK; }
int fdtd_enable_peer_access(int) { fdtd::set_error("no peers in the emulated build"); return FDTD_EUNSUPPORTED; }
int fdtd_ipc_export(const void *, void *) { fdtd::set_error("no IPC in the emulated build"); return FDTD_EUNSUPPORTED; }
int fdtd_ipc_open(const void *, void **) { fdtd::set_error("no IPC in the emulated build"); return FDTD_EUNSUPPORTED; }
int fdtd_ipc_close(void *) { return FDTD_OK; }
}
'''


CHAIN_STUBS = r'''
#include "fd2d_march.cuh"
namespace fdtd_march {
bool chain_supported(int, bool) { return false; }
int launch_march_chain(const MarchParams<float> &, int, int, int, cudaStream_t, int) { fdtd::set_error("the warp-chain passes are not part of the emulated build"); return FDTD_EUNSUPPORTED; }
void preload_chain() {}
}
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


ASM = [  # the kernels' inline PTX, statement by statement -> its emulation (tests/emu/cuda_runtime.h, namespace emu)
    (r'asm volatile\("cp\.async\.c[ag]\.shared\.global \[%0\], \[%1\], (\d+), %2;" ::"r"\((\w+)\), "l"\((\w+)\), "r"\((\w+)\) : "memory"\);',
     r'emu::cp_async(\2, \3, \1, \4);'),
    (r'asm volatile\("cp\.async\.commit_group;" ::: "memory"\);', 'emu::cp_async_commit();'),
    (r'asm volatile\("cp\.async\.wait_group %0;" ::"n"\((\w+)\) : "memory"\);', r'emu::cp_async_wait(\1);'),
    (r'asm volatile\("ld\.acquire\.sys\.global\.u64 %0, \[%1\];" : "=l"\((\w+)\) : "l"\((\w+)\) : "memory"\);', r'\1 = emu::ld_acquire(\2);'),
    (r'asm volatile\("st\.release\.sys\.global\.u64 \[%0\], %1;" ::"l"\((\w+)\), "l"\((\w+)\) : "memory"\);', r'emu::st_release(\1, \2);'),
    (r'asm volatile\("mov\.u64 %0, %%globaltimer;" : "=l"\((\w+)\)\);', r'\1 = emu::global_timer_ns();'),
]


def _asm_operands(stmt: str):
    """the C expressions of an ``asm volatile("..." : outputs : inputs : clobbers);`` statement, in operand order"""
    import re
    ops, pos = [], 0
    body = re.sub(r'"(?:[^"\\]|\\.)*"\s*(?=["\n])', "", stmt)       # drop the PTX text (adjacent string literals)
    for m in re.finditer(r'"[=+]?[rlfhd]"\s*\(', stmt):
        a0 = m.end() - 1
        ops.append(stmt[a0 + 1:_match(stmt, a0, "(", ")") - 1].strip())
    return ops


# the multi-line PTX blocks of fd2d_chain.cu (mbarrier pipeline, TMA), by the instruction they are built around
BLOCKS = [
    ("cp.async.bulk.tensor.2d", lambda o: "emu::tma_issue6(%s);" % ", ".join(o)),
    ("mbarrier.try_wait.parity", lambda o: "emu::mbar_wait(%s, %s);" % (o[0], o[1])),
    ("mbarrier.init", lambda o: "emu::mbar_init(%s, %s);" % (o[0], o[1])),
    ("mbarrier.arrive.shared::cta.b64", lambda o: "emu::mbar_arrive(%s, %s);" % (o[0], o[1])),
    ("fence.mbarrier_init", lambda o: ";"),
    ("st.shared.v2.f32 [%0+8]", lambda o: "emu::st_shared_v2(%s + 8, %s, %s);" % (o[0], o[1], o[2])),
    ("st.shared.v2.f32 [%0]", lambda o: "emu::st_shared_v2(%s, %s, %s);" % (o[0], o[1], o[2])),
]


def rewrite_asm_blocks(src: str) -> str:
    out, pos = "", 0
    while True:
        i = src.find("asm volatile(", pos)
        if i < 0:
            return out + src[pos:]
        a0 = src.index("(", i)
        a1 = _match(src, a0, "(", ")")
        end = src.index(";", a1) + 1
        stmt = src[i:end]
        for key, make in BLOCKS:
            if key in stmt:
                out += src[pos:i] + make(_asm_operands(stmt))
                break
        else:
            out += src[pos:end]                         # left for the single-statement patterns below
        pos = end


def rewrite_device_code(src: str) -> str:
    import re
    for pat, rep in ASM:
        src = re.sub(pat, rep, src)
    src = rewrite_asm_blocks(src)
    if re.search(r"\basm\b", src):
        raise RuntimeError("inline PTX without an emulation: " + re.search(r"\basm\b.*", src).group(0)[:120])
    # extern __shared__ T name[];  ->  the CTA's dynamic shared memory
    src = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w ]+?)\s+(\w+)\[\];",
                 r"\1 *const \2 = reinterpret_cast<\1 *>(emu::blk->smem);", src)
    return src


def build(*source_names: str) -> str:
    """-> path of the emulated shared object made of simulation_b200/csrc/<source_names>"""
    srcs = {n: open(os.path.join(CSRC, n)).read() for n in source_names}
    cuh = {n: open(os.path.join(CSRC, n)).read() for n in sorted(os.listdir(CSRC)) if n.endswith(".cuh")}
    hdr = open(os.path.join(HERE, "cuda_runtime.h")).read() + "".join(cuh.values()) \
        + open(os.path.join(ROOT, "include", "fdtd_b200.h")).read() + open(__file__).read()
    tag = hashlib.sha256(("".join(srcs.values()) + hdr).encode()).hexdigest()[:16]
    bdir = os.path.join(HERE, "_build")
    try:
        os.makedirs(bdir, exist_ok=True)
        if not os.access(bdir, os.W_OK):
            raise OSError("not writable")
    except OSError:                                      # a read-only checkout: build beside the system's temp files
        import tempfile
        bdir = os.path.join(tempfile.gettempdir(), "fdtd_b200_emu_build")
        os.makedirs(bdir, exist_ok=True)
    stem = "_".join(os.path.splitext(n)[0] for n in source_names)
    so = os.path.join(bdir, f"emu_{stem}_{tag}.so")
    if os.path.exists(so):
        return so
    stubs = STUBS if "fd2d_steps.cu" not in srcs else STUBS.replace(STUBS[STUBS.index("int launch_fourier"):STUBS.index("}\nextern")], "")
    assert "capi_misc.cu" not in srcs, "capi_misc.cu is replaced by the stubs"
    # the kernels' shared headers carry inline PTX too: rewritten copies shadow the originals on the include path
    inc = os.path.join(bdir, f"inc_{tag}")
    os.makedirs(inc, exist_ok=True)
    for n, text in cuh.items():
        with open(os.path.join(inc, n), "w") as f:
            f.write(rewrite_device_code(text).replace('"../../include/fdtd_b200.h"', '"fdtd_b200.h"'))
    if "fd2d_deep.cu" in srcs and "fd2d_chain.cu" not in srcs:      # a partial build without the warp-chain kernel
        stubs += CHAIN_STUBS
    units = {"stubs": '#include "common.cuh"\n' + stubs}
    for n, src in srcs.items():
        units[os.path.splitext(n)[0]] = f'#line 1 "{n}"\n' + rewrite_launches(rewrite_device_code(src))
    def compile_unit(item):
        name, text = item
        cpp = os.path.join(bdir, f"emu_{name}_{tag}.cpp")
        with open(cpp, "w") as f:
            f.write(text)
        obj = os.path.join(bdir, f"emu_{name}_{tag}.o")
        subprocess.run(["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-fno-strict-aliasing", "-fPIC", "-pthread", "-w",
                        "-I", inc, "-I", HERE, "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-x", "c++", "-c", cpp, "-o", obj], check=True)
        os.remove(cpp)
        return obj

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(len(units)) as pool:
        objs = list(pool.map(compile_unit, units.items()))
    subprocess.run(["g++", "-shared", "-pthread", *objs, "-o", so + ".tmp"], check=True)
    os.replace(so + ".tmp", so)
    for o in objs:
        os.remove(o)
    return so


ALL_SOURCES = ("fd1d.cu", "fd2d_steps.cu", "fd2d_march.cu", "fd2d_deep.cu", "fd2d_chain.cu")


def build_library() -> str:
    """the whole library (every kernel source; capi_misc.cu replaced by host stubs) as one emulated shared object"""
    return build(*ALL_SOURCES)


if __name__ == "__main__":
    import sys
    print(build(*(sys.argv[1:] or ["fd1d.cu"])))
