#!/usr/bin/env python
"""Build tools/cuemu/_build/libvorbis_b200_emu.so: the product's CUDA sources compiled for the HOST on
top of tools/cuemu/cuemu.h (CUDA threads = OS threads).  Development aid only - see cuemu.h.

The sources are copied to _build/src and two textual rewrites are applied to the copies:
  kernel<<<grid, block, smem, stream>>>(args);   ->  cuemu::launch(grid, block, smem, [&]{ kernel(args); });
  extern __shared__ [__align__(n)] T name[];     ->  T *name = (T *)cuemu::dyn_smem();
Everything else (qualifiers, intrinsics, the few runtime calls) is provided by cuemu.h.
"""
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "vorbis_b200", "csrc")
BUILD = os.path.join(HERE, "_build")
OUT = os.path.join(BUILD, "libvorbis_b200_emu.so")


def rewrite_launches(src):
    out = []
    i = 0
    while True:
        j = src.find("<<<", i)
        if j < 0:
            out.append(src[i:])
            break
        # kernel expression: identifier with optional template arguments, directly before <<<
        k = j
        depth = 0
        while k > i:
            c = src[k - 1]
            if c == ">":
                depth += 1
            elif c == "<":
                depth -= 1
            elif depth == 0 and not (c.isalnum() or c in "_:"):
                break
            k -= 1
        kernel = src[k:j]
        e = src.index(">>>", j)
        cfg = src[j + 3:e]
        # split the launch configuration at top-level commas
        parts, depth, cur = [], 0, ""
        for c in cfg:
            if c in "([{":
                depth += 1
            elif c in ")]}":
                depth -= 1
            if c == "," and depth == 0:
                parts.append(cur); cur = ""
            else:
                cur += c
        parts.append(cur)
        while len(parts) < 3:
            parts.append("0")
        # arguments: the parenthesised list after >>>
        a = e + 3
        while src[a].isspace() or src[a] == "\\":
            a += 1
        assert src[a] == "(", "launch without argument list near: " + src[j - 40:j + 80]
        depth, b = 0, a
        while True:
            if src[b] == "(":
                depth += 1
            elif src[b] == ")":
                depth -= 1
                if depth == 0:
                    break
            b += 1
        args = src[a:b + 1]
        out.append(src[i:k])
        out.append("cuemu::launch(dim3(%s), dim3(%s), (size_t)(%s), [&]() { %s%s; })"
                   % (parts[0].strip(), parts[1].strip(), parts[2].strip(), kernel, args))
        i = b + 1
    return "".join(out)


def rewrite(src):
    src = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w ]+?)\s+(\w+)\[\];",
                 r"\1 *\2 = (\1 *)cuemu::dyn_smem();", src)
    src = rewrite_launches(src)
    src = src.replace("#include <cuda_runtime.h>", '#include "cuemu.h"')
    return src


def build(verbose=True):
    srcdir = os.path.join(BUILD, "src")
    os.makedirs(srcdir, exist_ok=True)
    newest = 0.0
    for f in sorted(os.listdir(CSRC)):
        p = os.path.join(CSRC, f)
        newest = max(newest, os.path.getmtime(p))
        with open(p) as fh:
            text = fh.read()
        with open(os.path.join(srcdir, f), "w") as fh:
            fh.write(rewrite(text))
    for f in ("cuemu.h", "build_emu.py"):
        newest = max(newest, os.path.getmtime(os.path.join(HERE, f)))
    shutil.copy(os.path.join(HERE, "cuemu.h"), os.path.join(srcdir, "cuemu.h"))
    if os.path.exists(OUT) and os.path.getmtime(OUT) >= newest:
        return OUT
    glue = os.path.join(srcdir, "emu_main.cpp")
    with open(glue, "w") as fh:
        fh.write('#include "cuemu.h"\n'
                 "thread_local uint3 threadIdx, blockIdx;\nthread_local dim3 blockDim, gridDim;\n"
                 "namespace cuemu { thread_local ThreadCtx tctx; int g_sm_count = 2; }\n"
                 '#include "vb200.cu"\n')
    cmd = ["g++", "-std=c++20", "-O1", "-g", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-fno-fast-math",
           "-fno-strict-aliasing", "-DVB200_EMU=1", "-w", "-x", "c++", glue, "-I", srcdir,
           "-I", os.path.join(ROOT, "include"), "-o", OUT]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build())
