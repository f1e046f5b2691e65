"""Objective plugins: user objectives as device functors (include/nls_b200.h, nls_load_objective).

`compile_objective` turns the source of a functor (see nlsolver_b200/csrc/objective_plugin.cuh for the contract) into a
shared library with nvcc — the same toolchain that builds the engine — and `load_objective` registers it, returning the
objective id to put into `de_cfg(objective=...)` / `pso_cfg(objective=...)` or to pass as `f` to `DE` / `PSO`."""
import ctypes as C
import hashlib
import os
import subprocess

from . import _lib as L

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

TEMPLATE = """#include "objective_plugin.cuh"
{source}
NLS_EXPORT_OBJECTIVE({name})
"""


def headers_digest():
    """Content hash of every header a plugin compiles against (csrc/*.cuh, csrc/*.h, include/nls_b200.h): a cached
    plugin built from other kernel templates or state layouts is never reused."""
    h = hashlib.sha256()
    csrc = os.path.join(HERE, "csrc")
    names = sorted(n for n in os.listdir(csrc) if n.endswith((".cuh", ".h")))
    for path in [os.path.join(csrc, n) for n in names] + [os.path.join(ROOT, "include", "nls_b200.h")]:
        with open(path, "rb") as f:
            h.update(path.encode() + b"\0" + f.read())
    return h.hexdigest()


def compile_objective(source, name, out_dir=None, nvcc="nvcc", extra_flags=()):
    """source: C++ text defining `template <class T> struct <name>` with pairwise / lane0_seed / term / finish.
    extra_flags: e.g. ("-fmad=false",) to keep the functor free of FMA contraction (bit-reproducible against a host
    evaluation of the same expression).  Returns the path of the built shared library (cached by content hash)."""
    out_dir = out_dir or os.path.join(os.path.expanduser("~"), ".cache", "nlsolver_b200", "objectives")
    os.makedirs(out_dir, exist_ok=True)
    text = TEMPLATE.format(source=source, name=name)
    tag = hashlib.sha256((text + str(L.lib().nls_version()) + repr(tuple(extra_flags)) + headers_digest()).encode()
                         ).hexdigest()[:16]
    cu, so = os.path.join(out_dir, f"{name}_{tag}.cu"), os.path.join(out_dir, f"lib{name}_{tag}.so")
    if not os.path.exists(so):
        with open(cu, "w") as f:
            f.write(text)
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
               *extra_flags, "-I", os.path.join(HERE, "csrc"), "-I", os.path.join(ROOT, "include"), cu, "-o", so]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for objective plugin %s:\n%s" % (name, r.stderr))
    return so


def load_objective(path):
    """Register a built plugin; returns its objective id (>= 100)."""
    oid = L.i32()
    L.check(L.lib().nls_load_objective(os.fsencode(path), C.byref(oid)))
    return oid.value
