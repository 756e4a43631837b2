import json,sys
d=json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
print("value", round(d["value"]/1e9,2), "ms/step", round(d["ms_per_step"],2), "region_s", round(d["run"]["timed_region_s"],3))
r=d["roofline"]; print("roofline", round(r["frac"],4), "events", round(r["per_launch_events"]["frac"],4), [round(x*1e3,1) for x in r["per_launch_events"]["launch_ms_by_ply"]])
e=d["e2e"]; print("e2e", round(e["value"]/1e9,2), e["variant"], {k:(round(v["value"]/1e9,2), round(v["frac_of_link_ceiling"],3)) for k,v in e["variants"].items()})
print("ceiling", {k:(round(v,1) if isinstance(v,float) else v) for k,v in e["pcie_ceiling"].items() if k!="how"})
print("clocks", d["clocks"].get("sm_mhz"), d["clocks"].get("reasons"), "launches", d["gpu_launches"])
if d.get("cpu_baseline"): print("cpu", round(d["cpu_baseline"]["value"]), d["cpu_baseline"]["cores"])
for k,v in d["extra"].items(): print(" ", k, v if not isinstance(v, dict) else {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a not in ("note","sample","envs_by_len_moves_at_start")})
