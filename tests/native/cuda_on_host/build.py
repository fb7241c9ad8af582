"""TEST INFRASTRUCTURE: compile libcxxdes_b200/csrc with g++ against the CUDA stand-in of this directory.

The library's sources are used as they are, read from libcxxdes_b200/csrc at build time; three textual rewrites turn
what g++ cannot parse into calls of cuda_runtime.h (this directory):
  * `kernel<...><<<grid, block, shared, stream>>>(args)`  ->  `::cuda_host::launch(grid, block, shared, stream, kernel<...>, "kernel<...>")(args)`
  * `extern __shared__ __align__(16) unsigned char smem[];`  ->  a pointer to the block's shared-memory arena
  * `asm volatile("ld/st.shared/global ...")` (explicit loads and stores, predicated stores, %globaltimer)  ->  the
    same accesses in C++.  (The PTX event slot of the headline kernel sits behind `#if defined(__CUDA_ARCH__)`, which
    this build leaves undefined: its C++ twin runs, as in tests/native/mm1_lane_host.cpp.)
The result, tests/native/_build/cuda_on_host/libcxxdes_b200_on_host.so, exports the C ABI of include/cxxdes_b200.h and
is loaded by tests/test_kernels_on_host.py only.  It is not a CPU path of the product: nothing under libcxxdes_b200/
knows about it, it is not built by __graft_entry__.build's product step and never sits where load_library() looks."""
import hashlib
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent.parent
CSRC = ROOT / "libcxxdes_b200" / "csrc"
OUT_DIR = ROOT / "tests" / "native" / "_build" / "cuda_on_host"
OUT = OUT_DIR / "libcxxdes_b200_on_host.so"

_LAUNCH = re.compile(
    r"(?P<kernel>[A-Za-z_][\w:]*(?:\s*<[^;{}<>]*(?:<[^;{}<>]*>[^;{}<>]*)*>)?)\s*<<<(?P<cfg>.*?)>>>\s*\(", re.S)
_EXTERN_SHARED = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?unsigned char\s+(\w+)\[\];")


def _split_top(text, sep):
    """Split at `sep` outside parentheses, braces and string literals."""
    parts, depth, start, i, in_str = [], 0, 0, 0, False
    while i < len(text):
        c = text[i]
        if in_str:
            if c == "\\":
                i += 1
            elif c == '"':
                in_str = False
        elif c == '"':
            in_str = True
        elif c in "([{":
            depth += 1
        elif c in ")]}":
            depth -= 1
        elif c == sep and depth == 0:
            parts.append(text[start:i])
            start = i + 1
        i += 1
    parts.append(text[start:])
    return parts


def _operands(section):
    ops = []
    for item in _split_top(section, ","):
        item = item.strip()
        if not item:
            continue
        m = re.match(r'"([^"]*)"\s*\((.*)\)\s*$', item, re.S)
        if not m:
            raise ValueError(f"asm operand not understood: {item!r}")
        ops.append((m.group(1), m.group(2).strip()))
    return ops


_TYPES = {"u16": "uint16_t", "u32": "uint32_t", "u64": "uint64_t"}


def _translate_ptx(ptx, ops):
    """The handful of PTX statements the kernels spell out, as C++ on the operand expressions."""
    def op(token):
        token = token.strip()
        m = re.fullmatch(r"%(\d+)", token)
        if not m:
            raise ValueError(f"PTX operand not understood: {token!r}")
        return "(" + ops[int(m.group(1))][1] + ")"

    out = []
    body = ptx.replace("{\n", " ", 1) if ptx.lstrip().startswith("{") else ptx
    if ptx.lstrip().startswith("{"):
        body = body.rstrip()
        assert body.endswith("}"), ptx
        body = body[:-1]
    for stmt in body.split(";"):
        stmt = " ".join(stmt.split())
        if not stmt:
            continue
        guard = ""
        m = re.match(r"@(\w+)\s+(.*)", stmt)
        if m:
            guard, stmt = f"if (ptx_{m.group(1)}) ", m.group(2)
        if re.fullmatch(r"\.reg \.pred (\w+)", stmt):
            out.append(f"bool ptx_{stmt.split()[-1]} = false;")   # (PTX names must not shadow the operands' C++ names)
        elif (m := re.fullmatch(r"setp\.ne\.u32 (\w+), (%\d+), 0", stmt)):
            out.append(f"ptx_{m.group(1)} = {op(m.group(2))} != 0;")
        elif (m := re.fullmatch(r"mov\.u64 (%\d+), %%globaltimer", stmt)):
            out.append(f"{guard}{op(m.group(1))} = ::cuda_host::globaltimer();")
        elif (m := re.fullmatch(r"ld\.(shared|global)\.v2\.u32 \{(%\d+), (%\d+)\}, \[(%\d+)\]", stmt)):
            fn = "lds" if m.group(1) == "shared" else "ldg"
            out.append(f"{guard}{{ const uint2 v2__ = ::cuda_host::{fn}<uint2>({op(m.group(4))}); {op(m.group(2))} = v2__.x; {op(m.group(3))} = v2__.y; }}")
        elif (m := re.fullmatch(r"st\.(shared|global)\.v2\.u32 \[(%\d+)\], \{(%\d+), (%\d+)\}", stmt)):
            fn = "sts" if m.group(1) == "shared" else "stg"
            out.append(f"{guard}::cuda_host::{fn}<uint2>({op(m.group(2))}, uint2{{(unsigned){op(m.group(3))}, (unsigned){op(m.group(4))}}});")
        elif (m := re.fullmatch(r"ld\.(shared|global)\.(u16|u32|u64) (%\d+), \[(%\d+)\]", stmt)):
            fn = "lds" if m.group(1) == "shared" else "ldg"
            out.append(f"{guard}{op(m.group(3))} = ::cuda_host::{fn}<{_TYPES[m.group(2)]}>({op(m.group(4))});")
        elif (m := re.fullmatch(r"st\.(shared|global)\.(u16|u32|u64) \[(%\d+)\], (%\d+)", stmt)):
            fn = "sts" if m.group(1) == "shared" else "stg"
            out.append(f"{guard}::cuda_host::{fn}<{_TYPES[m.group(2)]}>({op(m.group(3))}, ({_TYPES[m.group(2)]}){op(m.group(4))});")
        else:
            raise ValueError(f"PTX statement not understood: {stmt!r}")
    return "do { " + " ".join(out) + " } while (0)"


def _rewrite_asm(text, name):
    out, pos = [], 0
    for m in re.finditer(r"\basm\s+volatile\s*\(", text):
        if m.start() < pos:
            continue
        depth, i, in_str = 1, m.end(), False
        while depth:
            c = text[i]
            if in_str:
                if c == "\\":
                    i += 1
                elif c == '"':
                    in_str = False
            elif c == '"':
                in_str = True
            elif c == "(":
                depth += 1
            elif c == ")":
                depth -= 1
            i += 1
        inner = text[m.end():i - 1]
        sections = _split_top(inner, ":")
        strings = re.findall(r'"((?:[^"\\]|\\.)*)"', sections[0])
        if not strings or re.sub(r'"(?:[^"\\]|\\.)*"', "", sections[0]).strip():
            continue        # assembled from macros (the PTX event slot): left alone, not compiled in this build
        ptx = "".join(strings).encode().decode("unicode_escape")
        ops = []
        for sec in sections[1:3]:
            ops += _operands(sec)
        try:
            replacement = _translate_ptx(ptx, ops)
        except ValueError as e:
            raise ValueError(f"{name}: {e}") from None
        out.append(text[pos:m.start()])
        out.append(replacement)
        pos = i
    out.append(text[pos:])
    return "".join(out)


def rewrite(text, name="<text>"):
    def launch(m):
        kernel = " ".join(m.group("kernel").split())
        return f'::cuda_host::launch({m.group("cfg").strip()}, {kernel}, "{kernel}")('
    text = _LAUNCH.sub(launch, text)
    text = _EXTERN_SHARED.sub(r"unsigned char *\1 = ::cuda_host::dynamic_shared();", text)
    return _rewrite_asm(text, name)


def _sources():
    return sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.h")) + sorted((CSRC / "device").glob("*"))


def _digest():
    h = hashlib.sha256()
    for p in _sources() + sorted(HERE.glob("*")) + sorted((ROOT / "include").rglob("*.h*")) + sorted((ROOT / "include" / "cxxdes_b200" / "shared").glob("*")):
        if p.is_file():
            h.update(p.name.encode())
            h.update(p.read_bytes())
    return h.hexdigest()


def build(force=False, verbose=False, asan=False):
    """Build (or reuse) the emulated library; returns its path.  asan: with -fsanitize=address, as a memory checker of
    the kernels (load it with LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0)."""
    out_dir = OUT_DIR.parent / "cuda_on_host_asan" if asan else OUT_DIR
    out = out_dir / OUT.name
    out_dir.mkdir(parents=True, exist_ok=True)
    stamp = out_dir / "digest.txt"
    digest = _digest()
    if not force and out.exists() and stamp.exists() and stamp.read_text() == digest:
        return out
    src_dir = out_dir / "src"
    (src_dir / "device").mkdir(parents=True, exist_ok=True)
    units = []
    for p in _sources():
        rel = p.relative_to(CSRC)
        target = src_dir / (rel.with_suffix(".cpp") if p.suffix == ".cu" else rel)
        target.write_text(rewrite(p.read_text(), str(rel)))
        if p.suffix == ".cu":
            units.append(target)
    units.append(HERE / "runtime.cpp")
    flags = ["-std=c++17", "-O1", "-g1", "-fPIC", "-ffp-contract=off", "-mfma", "-Wall", "-Wno-unknown-pragmas", "-Wno-unused-function",
             "-Wno-unused-variable", "-Wno-unused-but-set-variable", "-Wno-sign-compare", "-Wno-unused-local-typedefs",
             f"-I{HERE}", f"-I{ROOT / 'include'}", f"-I{src_dir}", "-include", "cuda_runtime.h"]
    if asan:
        flags += ["-fsanitize=address,undefined", "-fno-sanitize=vptr", "-fno-omit-frame-pointer"]   # memory errors and undefined behaviour (shifts, signed overflow, alignment)

    def compile_one(unit):
        obj = out_dir / (unit.stem + ".o")
        r = subprocess.run(["g++", *flags, "-c", str(unit), "-o", str(obj)], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"cuda_on_host: g++ failed on {unit.name}:\n{r.stderr[-6000:]}")
        if verbose and r.stderr:
            print(r.stderr, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=8) as pool:
        objs = list(pool.map(compile_one, units))
    subprocess.run(["g++", "-shared", *(["-fsanitize=address,undefined"] if asan else []), "-o", str(out), *map(str, objs), "-ldl"], check=True)
    # C++ hosts link `-lcxxdes_b200`: with LD_LIBRARY_PATH=<out_dir>/lib they find the emulated library under that name
    (out_dir / "lib").mkdir(exist_ok=True)
    alias = out_dir / "lib" / "libcxxdes_b200.so"
    if alias.is_symlink() or alias.exists():
        alias.unlink()
    alias.symlink_to(Path("..") / out.name)
    stamp.write_text(digest)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, asan="--asan" in sys.argv))
