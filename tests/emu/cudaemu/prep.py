#!/usr/bin/env python3
"""TEST INFRASTRUCTURE.  Copies necat_b200/csrc/*.cu|*.cuh|*.h into a build directory as C++ that g++ accepts, changing
only what is not C++:

  * `kernel<<<grid, block, smem, stream>>>(args)`  ->  `EMU_LAUNCH((kernel), grid, block, smem, stream)(args)`
  * `extern __shared__ [__align__(n)] T name[];`   ->  `T* name = (T*)emu::dyn_smem;`
  * the inline-PTX helpers (mbarrier / bulk copy / cp.async) are removed; cuda_runtime.h of this directory defines them;
  * the relative include of the public header becomes a plain one.

usage: prep.py <csrc dir> <out dir>
"""
import os
import re
import sys

PTX_HELPERS = ["nb_smem_addr", "nb_mbar_init", "nb_mbar_expect_tx", "nb_tma_load_1d", "nb_mbar_wait",
               "cns_cp_async16", "cns_cp_async_commit", "cns_cp_async_wait"]


def drop_function(src, name):
    m = re.search(r"^__device__ __forceinline__ [\w ]+\b%s\(" % re.escape(name), src, flags=re.M)
    if not m:
        return src
    i = src.index("{", m.end())
    depth, j = 0, i
    while True:
        c = src[j]
        if c == "{":
            depth += 1
        elif c == "}":
            depth -= 1
            if depth == 0:
                break
        j += 1
    return src[:m.start()] + "/* %s: see tests/emu/cudaemu/cuda_runtime.h */" % name + src[j + 1:]


def convert(src):
    n_launch = len(re.findall(r"<<<", src))
    src = re.sub(r"([A-Za-z_]\w*(?:<[^<>();]*>)?)\s*<<<([^;]*?)>>>\s*\(", r"EMU_LAUNCH((\1), \2)(", src)
    assert "<<<" not in src, "a kernel launch was not converted"
    src = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?(\w+)\s+(\w+)\[\];", r"\1* \2 = (\1*)emu::dyn_smem;", src)
    assert "extern __shared__" not in src
    for f in PTX_HELPERS:
        src = drop_function(src, f)
    src = src.replace('"../../include/necat_b200.h"', '"necat_b200.h"')          # found through -I include/
    assert "asm volatile" not in src and "asm(" not in src, "inline PTX left in the source"
    return src, n_launch


def strict_allocator(src):
    """nb_core.cu's caching allocator rounds sizes to 512 bytes and hands out larger cached blocks; with NB_EMU_STRICT=1
    in the environment the emulated build allocates exactly what was asked for, every time, so that AddressSanitizer sees
    an overrun of a single byte (tests/emu/README.md)."""
    a = "    bytes = (bytes + 511) & ~(size_t)511;\n"
    b = "        if (b.used || b.bytes < bytes || b.bytes > bytes * 2 + (8u << 20)) continue;\n"
    assert a in src and b in src, "nb_dev_alloc changed: update prep.py"
    src = src.replace(a, "    const bool emu_strict = getenv(\"NB_EMU_STRICT\") != NULL;\n    if (!emu_strict) bytes = (bytes + 511) & ~(size_t)511;\n")
    src = src.replace(b, "        if (emu_strict || b.used || b.bytes < bytes || b.bytes > bytes * 2 + (8u << 20)) continue;\n")
    return src


def main():
    csrc, out = sys.argv[1], sys.argv[2]
    os.makedirs(out, exist_ok=True)
    total = 0
    for f in sorted(os.listdir(csrc)):
        if not f.endswith((".cu", ".cuh", ".h")):
            continue
        text, n = convert(open(os.path.join(csrc, f)).read())
        if f == "nb_core.cu":
            text = strict_allocator(text)
        total += n
        name = f[:-3] + ".cpp" if f.endswith(".cu") else f
        open(os.path.join(out, name), "w").write(text)
    print("cudaemu prep: %d launches converted" % total)


if __name__ == "__main__":
    main()
