import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(round(d["value"]), round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["roofline"]["stage_ms"].items()}, round(d["roofline"]["frac"],4), d["roofline"]["kernel"])
