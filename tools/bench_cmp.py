#!/usr/bin/env python
"""Print the A/B bench lines written by tools/gpu_quick.sh."""
import json, sys
d0 = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/quick"
for n in ("tc", "mma"):
    try:
        d = json.load(open("%s/bench_%s.json" % (d0, n)))
        print(n, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step %.4f" % d["ms_per_step"], {k: v["ms"] for k, v in d["roofline"]["stages"].items()})
    except Exception as e:
        print(n, "ERR", e)
