"""Generate tests/golden/e2e_golden.json by running the UNMODIFIED reference HybridRetriever.retrieve + rerank
(reference src/advanced_rag/retrieval.py:215-339, 518-563) over the in-memory index manager -- test infrastructure.

Run in the build container only (needs /root/reference):  python -m oracle.gen_e2e_golden
"""
from __future__ import annotations

import json
import os

from . import e2e_corpus, inmem_index, ref_import

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "e2e_golden.json")


def main() -> None:
    ref_import.load()
    from advanced_rag.ranker import LearnedRanker
    from advanced_rag.retrieval import HybridRetriever, RetrievalConfig
    from advanced_rag.constants import RetrievalConstants
    RetrievalConstants.TIMEOUT_SECONDS = 60.0          # the CPU oracle is slow; the budget is not what is under test
    c = e2e_corpus.build()
    gen = inmem_index.HashEmbeddingGenerator(e2e_corpus.SEM_DIM, e2e_corpus.DOM_DIM, c["vocab"])
    mgr = inmem_index.InMemoryIndexManager(c["ids"], c["contents"], c["metadata"], c["semantic"], c["domain"],
                                           c["sp_ptr"], c["sp_idx"], c["sp_val"], e2e_corpus.VOCAB, gen)
    cases = []
    for text, kw in e2e_corpus.queries():
        r = HybridRetriever(mgr, RetrievalConfig(hybrid_alpha=0.7, top_k=20, enable_learned_ranker=True),
                            learned_ranker=LearnedRanker())
        hits = ref_import.run(r.retrieve(text, **kw))
        profile = hits[0]["metadata"]["retrieval_profile"] if hits else None
        rec = {"query": text, "kwargs": kw, "profile": profile,
               "ids": [h["id"] for h in hits], "scores_hex": [float(h["score"]).hex() for h in hits],
               "methods": [sorted(h["retrieval_methods"]) for h in hits],
               "method": [h["method"] for h in hits],
               "original_scores_hex": [float(h["original_score"]).hex() for h in hits]}
        # rerank is deterministic only where the learned ranker is on (default profile) or reranking is off (summary);
        # the other profiles fall into the reference's random-noise placeholder (retrieval.py:550-553)
        if r.config.enable_learned_ranker or not r.config.enable_reranking:
            rer = ref_import.run(r.rerank(text, [dict(h) for h in hits]))
            rec["rerank_ids"] = [h["id"] for h in rer]
            rec["rerank_scores_hex"] = [float(h["score"]).hex() for h in rer]
        cases.append(rec)
    with open(OUT, "w") as f:
        json.dump({"cases": cases, "n_docs": e2e_corpus.N_DOCS}, f, indent=1, sort_keys=True)
    print("wrote", OUT, len(cases), "queries;", sum(len(x["ids"]) for x in cases), "hits")


if __name__ == "__main__":
    main()
