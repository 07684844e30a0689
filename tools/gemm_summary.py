"""Condense tools/bench_gemm.py's JSON (stdin) to one line per shape: kernel -> (us, algorithmic GB/s)."""
import json
import sys

d = json.load(sys.stdin)
for k, v in d["gemm"].items():
    print(k, {m: (round(x["us"], 1), round(x["GBs"])) for m, x in v.items() if m != "tile"})
