import json,sys
tag=sys.argv[1]
d=json.loads(sys.stdin.read())
b=d["breakdown"]
print(tag, "val=%.0f ms/step=%.2f"%(d["value"],d["ms_per_step"]), " ".join("%s=%.2f"%(k[:12],v["ms"]) for k,v in b.items() if isinstance(v,dict) and v["ms"]>0.5))
