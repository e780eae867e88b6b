#!/bin/bash
# step time of one rank's share of the learning step on ONE GPU, for 2 / 4 / 8 ranks (a proxy for the scaling run)
for s in 2 4 8; do
  timeout 100 python bench.py --shard-of $s --steps 10 --warmup 3 --no-sweep --no-configs --no-cpu 2>/dev/null > /tmp/shard_$s.json
  python - "$s" <<'P'
import json, sys
s = int(sys.argv[1])
d = json.loads(open(f"/tmp/shard_{s}.json").read().strip().splitlines()[-1])
print("shard-of", s, "ms/step", round(d["ms_per_step"], 3), "ideal", round(47.09 / s, 3), "eff", round(47.09 / s / d["ms_per_step"], 3))
P
done
