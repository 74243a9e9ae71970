"""Print the device timeline bench.py --trace-file wrote: one line per launch with the idle gap before it."""
import json, sys
d = json.load(open(sys.argv[1]))
L = d["launches"]
print("wall ms", round(d["wall_ms"], 3), "launches", len(L))
end = None
busy = 0.0
for i, (c, t0, dt) in enumerate(L):
    gap = 0.0 if end is None else t0 - end
    busy += dt
    flag = "  <<<<" if gap > 0.02 else ""
    print(f"{i:4d} {c:9s} start {t0:8.3f} dur {dt*1e3:7.1f} us  gap {gap*1e3:7.1f} us{flag}")
    end = t0 + dt
print("span ms", round(end, 3), "busy ms", round(busy, 3), "idle ms", round(end - busy, 3))
