"""Collect the reference's 60 fault-target descriptors (input/encoder/*.json, input/decoder/*.json) into one fixture.
Run in the build container only (reads /root/reference); the fixture tests/golden/fault_targets.json travels."""
import glob
import json
import os

REF = "/root/reference/input"
out = {}
for module in ("encoder", "decoder"):
    rows = [json.load(open(f)) for f in sorted(glob.glob(os.path.join(REF, module, "*.json")))]
    rows.sort(key=lambda d: int(d["target_layer"].split("_")[1]))
    out[module] = rows
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fault_targets.json")
json.dump(out, open(path, "w"), indent=1)
print(path, {k: len(v) for k, v in out.items()})
