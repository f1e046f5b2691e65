"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol include/nls_b200.h declares,
and refuses to run without a CUDA device (no CPU fallback)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "nls_b200.h")).read()
    return sorted(set(re.findall(r"NLS_API [\w \*]*?\b(nls_\w+)\(", text)))


def test_library_exports_every_declared_symbol():
    from nlsolver_b200 import _lib
    handle = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/nls_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == names, "python binding table and header disagree"
    assert handle.nls_version() == 101


def test_struct_layouts_match_header_sizes():
    import ctypes as C
    from nlsolver_b200 import _lib
    assert C.sizeof(_lib.DECfg) == 96 and C.sizeof(_lib.PSOCfg) == 112 and C.sizeof(_lib.Status) == 88
    assert C.sizeof(_lib.SANNCfg) == 72
    assert _lib.lib().nls_record_bytes(_lib.F64, 256) == 48 + 256 * 8
    assert _lib.lib().nls_record_bytes(_lib.F32, 3) == 48 + 16


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import nlsolver_b200 as nb
    with pytest.raises(nb.NlsError) as e:
        nb.Context(0)
    assert "no CPU path" in str(e.value)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under nlsolver_b200/ or include/ may import, include, link or
    dlopen anything under oracle/ (comments may cite it)."""
    bad = re.compile(r"(import\s+oracle|from\s+oracle|#include\s+[\"<][^\">]*oracle|liboracle|libnls_ref|oracle_abi|"
                     r"oracle_de_run|oracle_pso_run|oracle_sann_run|ref_de_run|ref_pso_run|ref_sann_run)")
    for top in ("nlsolver_b200", "include"):
        for dirpath, dirs, files in os.walk(os.path.join(ROOT, top)):
            dirs[:] = [d for d in dirs if d not in ("build", "__pycache__")]
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", "Makefile")):
                    text = open(os.path.join(dirpath, f)).read()
                    assert not bad.search(text), f"{os.path.join(dirpath, f)} reaches into oracle/"


def test_seed_from_generator_consumes_two_draws():
    from nlsolver_b200.solvers import seed_from_generator
    draws = iter([0.5, 0.25, 0.9])
    seed = seed_from_generator(lambda: next(draws))
    assert seed == (0x80000000 << 32) | 0x40000000 and next(draws) == 0.9


def test_null_arguments_are_rejected_before_any_cuda_call():
    """Argument validation is host logic: it must answer NLS_ERR_INVALID (and leave a message) even without a GPU."""
    import ctypes as C
    from nlsolver_b200 import _lib
    h = _lib.lib()
    out = C.c_void_p()
    st = _lib.Status()
    calls = [
        lambda: h.nls_ctx_create(0, None, None),
        lambda: h.nls_de_create(None, None, None, C.byref(out)),
        lambda: h.nls_pso_create(None, None, None, None, C.byref(out)),
        lambda: h.nls_sann_create(None, None, None, 1, C.byref(out)),
        lambda: h.nls_de_solve(None, None, None, None, C.byref(st)),
        lambda: h.nls_pso_solve(None, None, None, None, None, C.byref(st)),
        lambda: h.nls_sann_solve(None, None, None, 1, None, C.byref(st)),
        lambda: h.nls_de_step(None, 1), lambda: h.nls_pso_step(None, 1), lambda: h.nls_sann_step(None, 1),
        lambda: h.nls_de_sync(None, None), lambda: h.nls_pso_sync(None, None), lambda: h.nls_sann_sync(None, None),
        lambda: h.nls_sann_read_best(None, None), lambda: h.nls_sann_read_chains(None, None, None, None, None, None),
        lambda: h.nls_load_objective(None, None),
        lambda: h.nls_xchg_create(None, 64, 1, 0, C.byref(out)),
    ]
    for k, call in enumerate(calls):
        assert call() == -1, f"call {k} did not return NLS_ERR_INVALID"
        assert h.nls_last_error(), f"call {k} left no message"
    # destroying nothing is fine, as free(NULL) is
    assert h.nls_de_destroy(None) == 0 and h.nls_pso_destroy(None) == 0 and h.nls_sann_destroy(None) == 0
    assert h.nls_ctx_destroy(None) == 0 and h.nls_xchg_destroy(None) == 0
    assert h.nls_load_objective(b"/nonexistent/libobjective.so", C.byref(C.c_int32())) == -1
    assert b"nonexistent" in h.nls_last_error()
