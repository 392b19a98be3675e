"""TEST INFRASTRUCTURE - NOT PRODUCT CODE. Builds tests/cpu_emul/_build/libopenifem_b200_cpuemul.so: the sources of
openifem_b200/csrc, mechanically rewritten so that g++ accepts them (kernel launches become calls into the SIMT emulator of
cpu_emul_engine.cpp, `extern __shared__` arrays point at the emulator's buffer, the eight inline-PTX loads become plain loads)
and compiled against the stand-in cuda_runtime.h of this directory. The kernel bodies, launch configurations and host
orchestration are the product's, unchanged. Only tests/cpu_emul/run_tests.py loads the result."""
from __future__ import annotations

import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SRC = os.path.join(ROOT, "openifem_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libopenifem_b200_cpuemul.so")
PRIMS = re.compile(r"__syncthreads|__syncwarp|__shfl_")
KEYWORDS = {"if", "for", "while", "switch", "catch", "return", "sizeof", "static_assert", "alignas", "decltype", "defined"}


def match(text, i, open_c, close_c):
    """index just past the bracket that closes text[i] (which must be open_c)"""
    depth = 0
    while i < len(text):
        c = text[i]
        if c == open_c:
            depth += 1
        elif c == close_c:
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1
    raise ValueError("unbalanced")


def function_bodies(text):
    """rough scan: name(...) [qualifiers] { body } -> {name: concatenated bodies}"""
    out = {}
    for m in re.finditer(r"\b([A-Za-z_]\w*)\s*\(", text):
        name = m.group(1)
        if name in KEYWORDS:
            continue
        try:
            j = match(text, m.end() - 1, "(", ")")
        except ValueError:
            continue
        k = j
        while True:
            mm = re.match(r"\s*(const|noexcept|override|final)\b", text[k:])
            if not mm:
                break
            k += mm.end()
        mm = re.match(r"\s*\{", text[k:])
        if not mm:
            continue
        try:
            e = match(text, k + mm.end() - 1, "{", "}")
        except ValueError:
            continue
        out[name] = out.get(name, "") + text[k:e]
    return out


def barrier_functions(texts):
    bodies = {}
    for t in texts:
        for n, b in function_bodies(t).items():
            bodies[n] = bodies.get(n, "") + b
    flagged = {n for n, b in bodies.items() if PRIMS.search(b)}
    changed = True
    while changed:
        changed = False
        pat = re.compile(r"\b(" + "|".join(map(re.escape, flagged)) + r")\b") if flagged else None
        for n, b in bodies.items():
            if n not in flagged and pat and pat.search(b):
                flagged.add(n)
                changed = True
    return flagged


def split_top(s):
    parts, depth, cur = [], 0, ""
    for c in s:
        if c in "([{<":
            depth += 1
        elif c in ")]}>":
            depth -= 1
        if c == "," and depth == 0:
            parts.append(cur)
            cur = ""
        else:
            cur += c
    parts.append(cur)
    return [p.strip() for p in parts]


def rewrite_launches(text, barriers, fname):
    out, pos = "", 0
    while True:
        i = text.find("<<<", pos)
        if i < 0:
            return out + text[pos:]
        # kernel expression: identifier, optionally followed by template arguments, right before <<<
        j = i
        while j > 0 and text[j - 1].isspace():
            j -= 1
        if text[j - 1] == ">":
            depth, k = 0, j - 1
            while True:
                if text[k] == ">":
                    depth += 1
                elif text[k] == "<":
                    depth -= 1
                    if depth == 0:
                        break
                k -= 1
            j = k
        k = j
        while k > 0 and (text[k - 1].isalnum() or text[k - 1] in "_:"):
            k -= 1
        kernel = text[k:i].strip()
        base = re.match(r"[\w:]+", kernel).group(0).split("::")[-1]
        e = text.index(">>>", i)
        cfg = split_top(text[i + 3:e])
        a0 = text.index("(", e)
        a1 = match(text, a0, "(", ")")
        args = text[a0 + 1:a1 - 1]
        grid, block = cfg[0], cfg[1]
        smem = cfg[2] if len(cfg) > 2 else "0"
        flag = "true" if base in barriers else "false"
        out += text[pos:k] + (f"cpu_emul::launch(dim3({grid}), dim3({block}), (size_t)({smem}), {flag}, \"{base} ({fname})\", "
                              f"[&]() {{ {kernel}({args}); }})")
        pos = a1


ASM_LOAD = re.compile(r'asm volatile\("ld\.global[^"]*"\s*:\s*"=\w"\((\w+)(?:\.\w+)?\)(?:\s*,\s*"=\w"\(\w+\.\w+\))*\s*:\s*"l"\((.*?)\)(?:\s*,\s*"l"\(\w+\))?\);')


def rewrite(text, barriers, fname):
    text = text.replace('#include "../../include/openifem_b200.h"', "#include <openifem_b200.h>")
    text = re.sub(r"extern\s+__shared__\s+(\w+)\s+(\w+)\[\];", r"\1 *\2 = reinterpret_cast<\1 *>(cpu_emul::dyn_smem());", text)
    # code that exists in two forms: `#ifdef IFEM_EMULATED_DEVICE <plain C++> #else <inline PTX> #endif` keeps the plain form
    text = re.sub(r"#ifdef IFEM_EMULATED_DEVICE\n(.*?)#else\n.*?#endif\n", r"\1", text, flags=re.S)
    text = ASM_LOAD.sub(lambda m: f"std::memcpy(&{m.group(1)}, (const void *)({m.group(2)}), sizeof({m.group(1)}));", text)
    text = re.sub(r'asm volatile\("createpolicy[^;]*;"\s*:\s*"=l"\((\w+)\)\);', r"\1 = 0;", text)
    if "asm volatile" in text:
        raise RuntimeError(f"{fname}: an inline-asm statement was not rewritten")
    return rewrite_launches(text, barriers, fname)


def build(force=False):
    os.makedirs(os.path.join(OUT, "src"), exist_ok=True)
    names = sorted(os.listdir(SRC))
    texts = {n: open(os.path.join(SRC, n)).read() for n in names}
    deps = [os.path.join(SRC, n) for n in names] + [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".h", ".cpp", ".py"))]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    barriers = barrier_functions(texts.values())
    units = []
    for n, t in texts.items():
        dst = os.path.join(OUT, "src", n[:-3] + ".cpp" if n.endswith(".cu") else n)
        new = rewrite(t, barriers, n)
        if not os.path.exists(dst) or open(dst).read() != new:
            open(dst, "w").write(new)
        if n.endswith((".cu", ".cpp")) and n != "comm.cpp":  # NCCL is replaced by the file-based communicator of comm_emul.cpp
            units.append(dst)
    units.append(os.path.join(HERE, "cpu_emul_engine.cpp"))
    units.append(os.path.join(HERE, "comm_emul.cpp"))
    flags = ["-std=c++17", "-O1", "-g", "-fPIC", "-fopenmp", "-ffp-contract=off", "-Wno-unknown-pragmas", "-Wno-deprecated", "-w",
             "-D__CUDACC__", "-DIFEM_EMULATED_DEVICE", "-include", os.path.join(HERE, "cuda_runtime.h"), "-I", HERE, "-I", os.path.join(OUT, "src"), "-I", os.path.join(ROOT, "include")]

    def compile_one(src):
        obj = os.path.join(OUT, os.path.basename(src) + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(d) for d in deps):
            r = subprocess.run(["g++"] + flags + ["-c", src, "-o", obj], capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"g++ failed on {src}:\n{r.stderr[-6000:]}")
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, units))
    subprocess.check_call(["g++", "-shared", "-o", LIB + ".tmp"] + objs + ["-fopenmp", "-ldl"])
    os.replace(LIB + ".tmp", LIB)  # a process that has the old file mapped keeps it
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
