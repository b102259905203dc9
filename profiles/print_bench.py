import json,sys
d=json.load(open(sys.argv[1]))
print(sys.argv[1], "step %.4f kernel %.4f" % (d["ms_per_step"], d["roofline"]["kernel_ms"]), " ".join("%s %.4f" % (a["workload"][:7], a["ms_per_step"]) for a in d["also"]))
