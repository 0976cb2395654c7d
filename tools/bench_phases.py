"""One-line digest of a bench.py JSON line: value, e2e, ms per step and the CUDA-event phase times.  usage: bench_phases.py line.json"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d.get("roofline", {})
print(f"{sys.argv[1]}: {d['value']/1e6:.2f} M rays/s, {d['ms_per_step']:.3f} ms/step, e2e {d.get('e2e', {}).get('value', 0)/1e6:.2f} M; "
      f"phases {r.get('phase_ms')}")
