"""Single-replan latency (bench.py's replan_latency) by engine mode: where the packed inputs / outputs live (device blocks, outputs
in mapped pinned memory, both) and whether K1 + evaluate + K5 + gather run as one kernel (FRENET_STAGE_ONE_LAUNCH)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
out = {}
modes = [("device_blocks,separate", dict(_zero_copy=False, _one_launch=False)),
         ("device_blocks,one_launch", dict(_zero_copy=False, _one_launch=True, _inline=False)),
         ("out_mapped,separate", dict(_zero_copy="out", _one_launch=False)),
         ("out_mapped,one_launch,graph (default)", dict(_zero_copy="out", _one_launch=True)),
         ("out_mapped,one_launch,inline", dict(_zero_copy="out", _one_launch=True, _inline=True)),
         ("out_mapped,one_launch,graph,no_speculation", dict(_zero_copy="out", _one_launch=True, _speculate=False)),
         ("all_mapped,one_launch", dict(_zero_copy=True, _one_launch=True)),
         ("round1", dict(_zero_copy=False, _one_launch=False, _speculate=False))]
for name, kw in modes:
    bench.replan_latency(40, **kw)
    out[name] = bench.replan_latency(400, **kw)
print(json.dumps(out))
for k in out[modes[0][0]]:
    print(k, {m: round(out[m][k]["p50_us"], 1) for m in out})
