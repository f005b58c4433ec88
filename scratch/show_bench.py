import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["n_gpus"], "ms", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d["config"].get("parallelism"),
              {k: (round(v["mean_ms"], 3), round(v["share_of_step"], 3)) for k, v in d.get("kernels", {}).items()}, "roofline", (d.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(f, "no json:", e)
