"""Generate tests/golden/pipeline_golden.json by running the UNMODIFIED reference AdvancedRAGPipeline.retrieve
(reference src/advanced_rag/pipeline.py:217-309: rewrite -> HybridRetriever.retrieve -> rerank(top_k=rerank_top_k) -> evaluate
-> audit -> RetrievalResult) over the in-memory index manager -- test infrastructure.  The fixture pins what the caller of the
pipeline gets: chunk ids, fused / re-ranked scores, retrieval method tags, metadata.

Run in the build container only (needs /root/reference):  python -m oracle.gen_pipeline_golden
"""
from __future__ import annotations

import json
import os

from . import e2e_corpus, inmem_index, ref_import

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "pipeline_golden.json")


def main() -> None:
    ref_import.load()
    from advanced_rag.constants import RetrievalConstants
    from advanced_rag.pipeline import AdvancedRAGPipeline, PipelineConfig
    from advanced_rag.ranker import LearnedRanker
    from advanced_rag.retrieval import HybridRetriever, RetrievalConfig
    RetrievalConstants.TIMEOUT_SECONDS = 60.0          # the CPU oracle is slow; the budget is not what is under test
    c = e2e_corpus.build()
    gen = inmem_index.HashEmbeddingGenerator(e2e_corpus.SEM_DIM, e2e_corpus.DOM_DIM, c["vocab"])
    mgr = inmem_index.InMemoryIndexManager(c["ids"], c["contents"], c["metadata"], c["semantic"], c["domain"],
                                           c["sp_ptr"], c["sp_idx"], c["sp_val"], e2e_corpus.VOCAB, gen)
    cfg = PipelineConfig()
    cfg.enable_audit_logging = False                   # the audit trail needs the compliance store (out of scope)
    cases = []
    for text, kw in e2e_corpus.queries():
        if kw.get("use_domain_index"):                 # the pipeline's retrieve never asks for the domain index (pipeline.py:240-246)
            continue
        pipe = AdvancedRAGPipeline(config=cfg, connect_to_milvus=False)
        # the two-line swap INTEGRATION.md shows: index manager + retriever; the learned ranker must be ENABLED for a
        # deterministic re-rank (the reference's default branch adds N(0, 0.01) noise, retrieval.py:550-553)
        pipe.index_manager = mgr
        pipe.retriever = HybridRetriever(mgr, RetrievalConfig(hybrid_alpha=cfg.hybrid_alpha, top_k=cfg.top_k,
                                                              enable_reranking=cfg.enable_reranking, enable_learned_ranker=True),
                                         learned_ranker=LearnedRanker())
        context = {"retrieval_profile": kw["profile_hint"]} if kw.get("profile_hint") else None
        results, metrics = ref_import.run(pipe.retrieve(text, context=context))
        learned = pipe.retriever.config.enable_learned_ranker or not pipe.retriever.config.enable_reranking
        if not learned:                                # profile without the learned ranker: random re-rank, nothing to pin
            continue
        cases.append({"query": text, "context": context, "rerank_top_k": cfg.rerank_top_k,
                      "chunk_ids": [r.chunk_id for r in results], "scores_hex": [float(r.score).hex() for r in results],
                      "retrieval_method": [r.retrieval_method for r in results], "content": [r.content for r in results],
                      "doc_ids": [r.metadata.get("doc_id") for r in results],
                      "profiles": [r.metadata.get("retrieval_profile") for r in results],
                      "n_results": len(results)})
    with open(OUT, "w") as f:
        json.dump({"cases": cases, "top_k": cfg.top_k, "rerank_top_k": cfg.rerank_top_k, "hybrid_alpha": cfg.hybrid_alpha,
                   "enable_reranking": cfg.enable_reranking}, f, indent=1, sort_keys=True)
    print("wrote", OUT, len(cases), "queries;", sum(x["n_results"] for x in cases), "results")


if __name__ == "__main__":
    main()
