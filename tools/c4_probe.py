import sys, json
sys.path[:0] = ["/root/repo", "/root/repo/advanced-rag-milvus_b200"]
import torch, bench_extras as bx
r = bx.c4("cuda:0")
print(json.dumps({k: v for k, v in r.items() if k != "workload"}))
