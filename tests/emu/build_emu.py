"""TEST INFRASTRUCTURE: builds tests/emu/_build/libgas_b200_emu.so — the product's .cu files (godot-audio-spatializer_b200/csrc,
byte for byte, except that `kernel<<<grid, block, smem, stream>>>(args)` is rewritten to a function call, which g++ cannot parse
otherwise) compiled by g++ against the stand-in CUDA runtime of this directory (cuda_runtime.h, emu_core.cpp, gas_ptx_emu.h).

    python tests/emu/build_emu.py            # incremental
    from tests.emu import build_emu; build_emu.build()  -> path of the library

The library exports the same C ABI as libgas_b200.so and runs every kernel on the CPU, one fiber per CUDA thread.  It is used by
tests only (tests/conftest.py, GAS_EMU=1); the product never loads it.
"""
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "godot-audio-spatializer_b200", "csrc")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libgas_b200_emu.so")

SOURCES = ["gas_api.cu", "gas_state.cu", "gas_gain.cu", "gas_prologue.cu", "gas_mix_stream.cu", "gas_mix_voice.cu", "gas_comm.cu",
           "gas_life.cu", "gas_single.cu", "gas_resample.cu", "gas_bus.cu"]

# kernel<<<grid, block, smem, stream>>>(args...)  ->  emu::launch("kernel", dim3(grid), dim3(block), smem, stream, kernel, args...)


def _split_top(s):
    """Split at top-level commas."""
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


def transform(text, name):
    """Rewrite every triple-chevron launch (possibly spanning lines; line breaks inside the argument list are kept)."""
    out, pos = "", 0
    head = re.compile(r"(?P<k>[A-Za-z_][A-Za-z0-9_]*(?:<[^<>;]*>)?)<<<(?P<cfg>.*?)>>>\(")
    while True:
        m = head.search(text, pos)
        if not m:
            if "<<<" in text[pos:]:
                raise RuntimeError(f"{name}: cannot rewrite a launch after offset {pos}")
            out += text[pos:]
            break
        depth, i = 1, m.end()
        while depth > 0:
            if i >= len(text):
                raise RuntimeError(f"{name}: unbalanced launch arguments")
            depth += text[i] == "("
            depth -= text[i] == ")"
            i += 1
        args = text[m.end():i - 1]
        cfg = _split_top(m.group("cfg"))
        if len(cfg) != 4:
            raise RuntimeError(f"{name}: expected <<<grid, block, smem, stream>>>: {m.group(0)}")
        k = m.group("k")
        out += text[pos:m.start()]
        out += f'emu::launch("{k.split("<")[0]}", dim3({cfg[0]}), dim3({cfg[1]}), {cfg[2]}, {cfg[3]}, {k}{", " + args if args.strip() else ""})'
        pos = i
    return out


def _newer(target, deps):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def build(verbose=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hdrs += [os.path.join(HERE, f) for f in ("cuda_runtime.h", "gas_ptx_emu.h")]
    hdrs += [os.path.join(ROOT, "include", "gas.h"), os.path.abspath(__file__)]
    cxx = os.environ.get("CXX", "g++")
    flags = ["-std=c++17", "-O2", "-g1", "-fPIC", "-fvisibility=hidden", "-ffp-contract=off", "-fno-strict-aliasing", "-pthread", "-w",
             "-I" + HERE, "-I" + CSRC, "-I" + os.path.join(ROOT, "include")]
    jobs = []
    objs = []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        gen = os.path.join(OUT_DIR, src.replace(".cu", ".emu.cpp"))
        obj = gen.replace(".cpp", ".o")
        objs.append(obj)
        text = transform(open(path).read(), src)
        text = f'#line 1 "{path}"\n' + text
        if not os.path.exists(gen) or open(gen).read() != text:
            open(gen, "w").write(text)
        if not _newer(obj, [gen] + hdrs):
            jobs.append([cxx] + flags + ["-c", gen, "-o", obj])
    core = os.path.join(HERE, "emu_core.cpp")
    core_obj = os.path.join(OUT_DIR, "emu_core.o")
    objs.append(core_obj)
    if not _newer(core_obj, [core] + hdrs):
        jobs.append([cxx] + flags + ["-c", core, "-o", core_obj])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("emulation build failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        if verbose:
            print(" ".join(cmd[-3:]))

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        list(ex.map(run, jobs))
    if jobs or not os.path.exists(LIB):
        run([cxx, "-shared", "-pthread", "-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    print(build(verbose=True))
