"""Import the UNMODIFIED reference package in the build container -- TEST INFRASTRUCTURE ONLY.

/root/reference exists only here (never on the GPU box), so nothing under `-m gpu`, smoke() or bench.py
may call this.  It is used by oracle/gen_golden.py to execute the reference's own fusion / MMR / rerank
code and by the CPU-only cross-check tests (skipped when the reference tree is absent).

The reference imports pymilvus at module top (src/advanced_rag/indexing.py:34-41), which is not
installed; a six-name stub satisfies the import.  No Milvus call is ever made.
"""
from __future__ import annotations

import asyncio
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("B200RAG_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "advanced_rag"))


def load():
    """Returns the imported `advanced_rag` package (reference code, unmodified)."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if "pymilvus" not in sys.modules:
        stub = types.ModuleType("pymilvus")
        for name in ("connections", "Collection", "CollectionSchema", "FieldSchema", "DataType", "utility"):
            setattr(stub, name, type(name, (), {}))
        sys.modules["pymilvus"] = stub
    src = os.path.join(REFERENCE_ROOT, "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    import advanced_rag  # noqa: E402
    return advanced_rag


def run(coro):
    """Run a coroutine on a fresh event loop (pytest-asyncio is not installed)."""
    loop = asyncio.new_event_loop()
    try:
        return loop.run_until_complete(coro)
    finally:
        loop.close()
