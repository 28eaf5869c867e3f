"""stdin: bench.py output; prints value + per-kernel ms of every JSON line (tuning runs)."""
import json
import sys
for line in sys.stdin:
    line = line.rstrip()
    if line.startswith("{"):
        d = json.loads(line)
        print(d["value"], d["kernel_ms_per_step"])
    elif line:
        print(line)
