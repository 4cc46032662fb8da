import json,sys
l=open(sys.argv[1]).read().strip().splitlines()[-1]
try:
    j=json.loads(l); print("ms/step %.4f  value %.3e  overflows %s"%(j["ms_per_step"], j["value"], j["config"]["table_overflows"])); print({k:round(v,4) for k,v in j["roofline_step"]["phases_ms"].items()})
except Exception as e: print(l[-800:])
