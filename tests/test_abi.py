"""CPU tests of the drop-in boundary: the shared library builds, loads, and exports every symbol the header
declares (no compute calls -- there is no GPU here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "b200rag.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200rag_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    import importlib.util
    spec = importlib.util.spec_from_file_location("b200rag_build", os.path.join(ROOT, "advanced-rag-milvus_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    from b200rag import _lib
    lib = _lib.load()
    syms = _header_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/b200rag.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == syms
    assert lib.b200rag_abi_version() == 2


def test_options_are_explicit_and_reentrant_from_two_threads():
    """reentrant, no global mutable state read from the environment.  A/B knobs are set through
    b200rag_set_option (atomics), the last-error string and the debug buffers are per thread; argument validation can run
    from two threads at once without interfering."""
    import threading
    from b200rag import _lib
    lib = _lib.load()
    assert lib.b200rag_get_option(b"sparse_slices") == -1 and lib.b200rag_get_option(b"no_such_option") == -2
    _lib.set_option("sparse_slices", 4)
    assert lib.b200rag_get_option(b"sparse_slices") == 4
    _lib.set_option("sparse_slices", -1)
    assert lib.b200rag_get_option(b"sparse_slices") == -1
    assert lib.b200rag_set_option(b"bogus", 1) == _lib.E_INVALID
    errors = {}

    def worker(tag, bad_dim):
        for _ in range(200):
            rc = lib.b200rag_prepare_rows(None, None, 4, bad_dim, 0, 0, None)
            msg = lib.b200rag_last_error().decode()
            if rc != _lib.E_INVALID or f"dim={bad_dim}" not in msg:
                errors[tag] = (rc, msg)
                return
            assert lib.b200rag_debug_set_stats_buffer(1, None, 0) == 0

    ts = [threading.Thread(target=worker, args=("a", -3)), threading.Thread(target=worker, args=("b", 9000))]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors


def test_fails_loudly_without_gpu_or_with_bad_arguments():
    import torch
    from b200rag import _lib, engine
    lib = _lib.load()
    # argument validation happens before any CUDA call
    rc = lib.b200rag_dense_topk(None, 10, 7, 0, None, 1, 1, 0, None, None, None, 1.0, None, None, 0, 1, None)
    assert rc == _lib.E_INVALID and b"null" in lib.b200rag_last_error()
    rc = lib.b200rag_sparse_topk(None, None, None, 1, 1, 33, None, None, None, 1, 1, 0, None, None, None, None, 0, None)
    assert rc == _lib.E_INVALID
    # no CPU path: host tensors are refused
    with pytest.raises(ValueError):
        engine.dense_topk(torch.zeros(4, 8, dtype=torch.float16), torch.zeros(1, 8, dtype=torch.float16), 1)
    with pytest.raises(ValueError):
        engine.DenseIndex(8, device="cpu")
    if not torch.cuda.is_available():
        import ctypes
        rc = lib.b200rag_device_info(None, None, None)
        assert rc == _lib.E_CUDA
        with pytest.raises(_lib.B200RagError):
            _lib.check(rc)
