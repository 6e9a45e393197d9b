"""Checker for `run_configs.py --configs 5` (not collected by pytest): recomputes, with the CPU oracle, the exact integer sums of the
candidates the run reported over the full global record axis and compares without tolerance.

    python -m torch.distributed.run ... run_configs.py --configs 5 > cfg5.jsonl
    python tests/check_config5.py cfg5.jsonl            # ~300 M oracle evaluations: seconds on 16 threads
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402


def main(path):
    oracle.build()
    ok = True
    for line in open(path):
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        if d.get("config") != 5:
            continue
        chk = d["check"]
        threads = min(16, len(os.sched_getaffinity(0)))
        es, en = oracle.score_generated_fx(d["seed"], chk["candidates"], 0, d["T_global"], 300, nthreads=threads)
        exact = [str(a) for a in es] == chk["sums"] and list(en) == chk["counts"]
        ok &= exact and chk["topk_follows_from_exact_sums"]
        print(json.dumps({"config": 5, "n_gpus": d["n_gpus"], "candidates": chk["candidates"], "records_per_candidate": d["T_global"],
                          "partials_exact_full_axis": exact, "topk_follows_from_exact_sums": chk["topk_follows_from_exact_sums"]}))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main(sys.argv[1]))
