#!/usr/bin/env python
"""Print `export` lines selecting the fastest exchange schedule of a shard_sweep.py output (argv[1]) for the 5-layer proposal."""
import json
import sys

best = None
for line in open(sys.argv[1]):
    try:
        d = json.loads(line)
    except Exception:
        continue
    if d.get("layers") == 5 and "mode" in d and (best is None or d["seconds_per_proposal"] < best["seconds_per_proposal"]):
        best = d
if best:
    print("export QUMCMC_SHARD_MODE=%s QUMCMC_SHARD_CHUNK_BITS=%d QUMCMC_SHARD_CTAS=%d" % (
        best["mode"], best.get("chunk_bits") or 0, best.get("xchg_ctas") or 148))
