#!/usr/bin/env python
"""Developer tool: copies the bench JSON lines of a round from gpurun_out/ into profiles/ and prints the table rows used in
DESIGN.md / profiles/r02_summary.md.   python tools/summarize_profiles.py name=path ..."""
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rows = []
    for arg in sys.argv[1:]:
        name, path = arg.split("=")
        try:
            d = json.loads(open(path).read().strip().split("\n")[-1])
        except Exception as e:
            print("skip", path, e)
            continue
        shutil.copy(path, os.path.join(ROOT, "profiles", name + ".json"))
        c = d.get("config", {})
        r = d.get("roofline") or {}
        rows.append("| %s | %d | %s | %.4f | %.3e | %.4f | %.3e | %s | %s |" % (
            name, d.get("n_gpus", 0), c.get("agents_total"), d.get("ms_per_step", 0) or 0, d.get("value", 0),
            (d.get("e2e", {}).get("ms_per_step") or 0), d.get("e2e", {}).get("value", 0),
            ("%.3f / %.3f" % (r.get("frac", 0), r.get("this_design", {}).get("frac", 0))) if r else "-",
            c.get("sharded_equals_single_gpu")))
    print("| line | GPUs | agents | ms per day (resident) | agent-days/s | ms per day (e2e) | e2e agent-days/s | roofline frac 8(d) / this design | sharded == single |")
    print("|---|---|---|---|---|---|---|---|---|")
    print("\n".join(rows))


if __name__ == "__main__":
    main()
