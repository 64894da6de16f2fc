import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("main", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d.get("parity"))
for k in ("north_star_c2","north_star_c4"):
    ns=d.get(k)
    if ns: print(k, {x:ns[x] for x in ("workload","value","ms_per_step","recall_at_10","recall_by_oversample","build_s","optimistic_reruns","stage_ms_per_step_rank0")}, "parity", ns["parity"] and (ns["parity"]["topk_ids_bit_exact"], ns["parity"]["scores_bit_exact"]), "e2e", ns["e2e"]["value"])
