import sys, importlib, json
sys.path[:0] = [".", "oracle"]
from gf2_loader import load
import bench
gf2 = load(); synth = importlib.import_module("gf2_b200.synth")
print(json.dumps(bench.run_replay(gf2, synth, with_cpu=False)))
