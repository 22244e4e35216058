import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print({k:d.get(k) for k in ("metric","value","n_gpus","ms_per_step","gpu_launches")}, "roofline", d["roofline"]["achieved"], d["roofline"]["frac"], "e2e", (d.get("e2e") or {}).get("value"))
except Exception as e:
    print("no json:", e, open(sys.argv[1]).read()[-600:])
