"""The small seeded hybrid corpus behind tests/golden/e2e_golden.json -- TEST INFRASTRUCTURE.

Pure numpy / Python so that the same inputs can be rebuilt bit for bit in the build container (where the reference
runs) and on the GPU box (where the CUDA path runs).  BM25 document weights are restated here independently of the
product's b200rag/bm25.py, as a plain per-posting fp64 loop of the formula DESIGN.md states:
    idf(t) = ln(1 + (N - df + 0.5) / (df + 0.5));  w = idf * tf * (k1 + 1) / (tf + k1 * (1 - b + b * len / avgdl)),
    k1 = 1.2, b = 0.75, rounded once to fp32.
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, Tuple

import numpy as np

N_DOCS, SEM_DIM, DOM_DIM, VOCAB = 3000, 64, 32, 400


def build(seed: int = 7) -> Dict[str, Any]:
    rng = np.random.default_rng(seed)
    p = 1.0 / np.arange(1, VOCAB + 1) ** 1.07
    p /= p.sum()
    ids, contents, metadata = [], [], []
    tf_rows: List[Dict[int, int]] = []
    vocab = {f"w{t}": t for t in range(VOCAB)}
    for d in range(N_DOCS):
        n_tok = int(rng.poisson(24)) if d % 97 else 0              # a few empty documents
        toks = rng.choice(VOCAB, size=n_tok, p=p)
        words = [f"w{t}" for t in toks]
        if words and d % 5 == 0:
            words[0] = words[0].upper()                              # exercises .lower()
        ids.append(f"c{d:06d}")
        contents.append(" ".join(words))
        metadata.append({"doc_id": f"d{d // 4}", "chunk_index": d % 4, "entropy": float(d % 10) / 10.0,
                         "redundancy": float(d % 7) / 7.0, "domain_density": float(d % 3) / 3.0,
                         "timestamp": f"2024-0{1 + d % 9}-15T00:00:00", "token_count": n_tok})
        cnt: Dict[int, int] = {}
        for t in toks:
            cnt[int(t)] = cnt.get(int(t), 0) + 1
        tf_rows.append(cnt)
    semantic = rng.standard_normal((N_DOCS, SEM_DIM)).astype(np.float32)
    domain = rng.standard_normal((N_DOCS, DOM_DIM)).astype(np.float32)
    semantic[11] = semantic[10]                                      # exact duplicates -> score ties broken by row
    semantic[12] = semantic[10]
    # BM25 (independent restatement)
    df = np.zeros(VOCAB, dtype=np.int64)
    for cnt in tf_rows:
        for t in cnt:
            df[t] += 1
    lens = [sum(c.values()) for c in tf_rows]
    avgdl = sum(lens) / N_DOCS
    sp_ptr, sp_idx, sp_val = [0], [], []
    for d, cnt in enumerate(tf_rows):
        for t in sorted(cnt):
            tf = float(cnt[t])
            idf = math.log(1.0 + (N_DOCS - float(df[t]) + 0.5) / (float(df[t]) + 0.5))
            w = idf * tf * (1.2 + 1.0) / (tf + 1.2 * (1.0 - 0.75 + 0.75 * float(lens[d]) / avgdl))
            sp_idx.append(t)
            sp_val.append(np.float32(w))
        sp_ptr.append(len(sp_idx))
    sparse_rows = [{"indices": sp_idx[sp_ptr[d]:sp_ptr[d + 1]], "values": [float(v) for v in sp_val[sp_ptr[d]:sp_ptr[d + 1]]]}
                   for d in range(N_DOCS)]
    return {"ids": ids, "contents": contents, "metadata": metadata, "semantic": semantic, "domain": domain,
            "sp_ptr": np.asarray(sp_ptr, dtype=np.int64), "sp_idx": np.asarray(sp_idx, dtype=np.int64),
            "sp_val": np.asarray(sp_val, dtype=np.float32), "sparse_rows": sparse_rows, "vocab": vocab,
            "tf_rows": tf_rows}


def queries() -> List[Tuple[str, Dict[str, Any]]]:
    """(query text, retrieve kwargs).  The texts hit every profile of the classifier."""
    rng = np.random.default_rng(99)
    out: List[Tuple[str, Dict[str, Any]]] = []

    def words(n):
        return " ".join(f"w{int(t)}" for t in rng.integers(3, 200, size=n))

    for i in range(6):
        out.append((words(6), {}))                                                     # default
    for i in range(4):
        out.append((words(4) + "?", {}))                                               # faq (top_k 10)
    for i in range(4):
        out.append(("error " + words(6), {}))                                          # troubleshooting (MMR 0.5, top_k 30)
    for i in range(3):
        out.append(("summary " + words(5), {}))                                        # summary (top_k 40)
    for i in range(3):
        out.append((words(45), {}))                                                    # analysis (>= 200 chars, MMR 0.8)
    for i in range(3):
        out.append((words(6), {"use_domain_index": True, "domain": "legal"}))          # three-way fusion
    out.append(("zzz qqq", {}))                                                        # no known token: sparse list empty
    out.append((words(5), {"profile_hint": "analysis"}))
    return out
