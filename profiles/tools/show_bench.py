import json,sys
d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d["value"]), d["ms_per_step"])
print("  "+" ".join(f"{k}={v['ms']}" for k,v in d["roofline"]["kernels"].items() if v["ms"]>0.3))
