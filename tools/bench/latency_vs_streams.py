import json, sys
sys.path.insert(0, '.')
import bench
from conan_b200 import synth
sds = synth.make_all_state_dicts(1234)
print(json.dumps(bench.leg_config1(0, sds)))
for S in (8, 64, 256):
    rig = bench.Rig(S, 0, sds)
    r = bench._short_window(rig, 50, 0.5)
    print(S, r["ms_per_step"], r["latency_ms"])
    rig.close()
