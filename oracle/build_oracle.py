"""Build recipe for the parity oracle (test infrastructure, never product code).

  liboracle.so            <- oracle/krep_oracle.c          (the CPU restatement; always built)
  _ref/libkrep_ref.so     <- /root/reference/{krep.c,aho_corasick.c} compiled where they lie,
                             through oracle/ref_wrap.c (adds accessors for krep.c's static flags)
  _ref/krep               <- the stock reference CLI, same sources, with its own main()
  _ref/libkrep_ref512.so  <- the same library as its AVX-512 build (adds simd_avx512_search)
  _ref/libkrep_refneon.so <- the same library as its NEON build (adds neon_search), compiled on x86 against
                             oracle/neon_shim/arm_neon.h (scalar stand-ins for the five intrinsics it uses)

The reference's own Makefile is NOT run; its flag set (Makefile:9-41) is restated here:
-O3 -std=c11 -pthread -D_GNU_SOURCE -D_DEFAULT_SOURCE -funroll-loops, SIMD flags fixed to
-msse4.2 -mavx2 (the AVX2 build; AVX-512 is left out so the binary also runs on GPU-box hosts
without it — the 8/16-byte literals of the BASELINE configs take the SSE4.2 kernel either way,
krep.c:4892).  -flto is dropped so every kernel keeps its symbol.

_ref/ is git-ignored; it is only (re)built when /root/reference exists (this container) and
travels to the GPU box as built files.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.environ.get("KREP_REF_DIR", "/root/reference")
OUT_REF = os.path.join(HERE, "_ref")
CFLAGS = ["-O3", "-std=c11", "-pthread", "-D_GNU_SOURCE", "-D_DEFAULT_SOURCE", "-funroll-loops",
          "-msse4.2", "-mavx2", "-w"]


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("oracle build failed")


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources if os.path.exists(s))


def build_port(force=False):
    src = os.path.join(HERE, "krep_oracle.c")
    hdr = os.path.join(HERE, "..", "include", "krep_b200.h")
    out = os.path.join(HERE, "liboracle.so")
    if force or _stale(out, [src, hdr]):
        _run(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-Wall", "-o", out, src])
    return out


def ref_available():
    return os.path.isfile(os.path.join(REF_DIR, "krep.c"))


def build_ref(force=False):
    """Returns (lib, cli) paths, or (None, None) when neither sources nor prebuilt files exist."""
    lib = os.path.join(OUT_REF, "libkrep_ref.so")
    cli = os.path.join(OUT_REF, "krep")
    if not ref_available():
        return (lib if os.path.exists(lib) else None, cli if os.path.exists(cli) else None)
    os.makedirs(OUT_REF, exist_ok=True)
    srcs = [os.path.join(REF_DIR, f) for f in ("krep.c", "aho_corasick.c", "krep.h", "aho_corasick.h")]
    wrap = os.path.join(HERE, "ref_wrap.c")
    if force or _stale(lib, srcs + [wrap]):
        _run(["gcc", *CFLAGS, "-DTESTING", "-fPIC", "-shared", "-I", REF_DIR, "-o", lib,
              wrap, os.path.join(REF_DIR, "aho_corasick.c")])
    if force or _stale(cli, srcs):
        _run(["gcc", *CFLAGS, "-I", REF_DIR, "-o", cli,
              os.path.join(REF_DIR, "krep.c"), os.path.join(REF_DIR, "aho_corasick.c")])
    return lib, cli


def build_ref512(force=False):
    """The AVX-512 build of the reference library (Makefile:34-35 flag set) — the only build in which
    simd_avx512_search exists.  Used by tests to pin oracle_avx512_search; never loaded on a CPU without AVX-512BW."""
    lib = os.path.join(OUT_REF, "libkrep_ref512.so")
    if not ref_available():
        return lib if os.path.exists(lib) else None
    os.makedirs(OUT_REF, exist_ok=True)
    srcs = [os.path.join(REF_DIR, f) for f in ("krep.c", "aho_corasick.c", "krep.h", "aho_corasick.h")]
    wrap = os.path.join(HERE, "ref_wrap.c")
    if force or _stale(lib, srcs + [wrap]):
        _run(["gcc", *CFLAGS, "-mavx512f", "-mavx512bw", "-DTESTING", "-fPIC", "-shared", "-I", REF_DIR, "-o", lib,
              wrap, os.path.join(REF_DIR, "aho_corasick.c")])
    return lib


def build_refneon(force=False):
    """The reference's ARM path (neon_search, krep.c:4506) compiled here: no x86 SIMD flags, -D__ARM_NEON, and
    oracle/neon_shim/arm_neon.h standing in for the five intrinsics it uses.  Pins oracle_neon_search."""
    lib = os.path.join(OUT_REF, "libkrep_refneon.so")
    if not ref_available():
        return lib if os.path.exists(lib) else None
    os.makedirs(OUT_REF, exist_ok=True)
    srcs = [os.path.join(REF_DIR, f) for f in ("krep.c", "aho_corasick.c", "krep.h", "aho_corasick.h")]
    wrap = os.path.join(HERE, "ref_wrap.c")
    shim = os.path.join(HERE, "neon_shim")
    if force or _stale(lib, srcs + [wrap, os.path.join(shim, "arm_neon.h")]):
        flags = [f for f in CFLAGS if f not in ("-msse4.2", "-mavx2")]
        _run(["gcc", *flags, "-mno-sse4.2", "-mno-avx", "-D__ARM_NEON", "-DTESTING", "-fPIC", "-shared", "-I", shim, "-I", REF_DIR,
              "-o", lib, wrap, os.path.join(REF_DIR, "aho_corasick.c")])
    return lib


if __name__ == "__main__":
    print(build_port(force="--force" in sys.argv))
    print(build_ref(force="--force" in sys.argv))
    print(build_ref512(force="--force" in sys.argv))
    print(build_refneon(force="--force" in sys.argv))
