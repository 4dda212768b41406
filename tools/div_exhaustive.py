"""Exhaustive check of the hoisted exact division (csrc/device_math.cuh: div_prepare + div_hoisted and its packed form)
against the compiler's IEEE division: ALL 2^23 x 2^23 mantissa pairs at one exponent / sign pair (task 5c).
usage: python tools/div_exhaustive.py [out.json] [ex_x ex_d signs] [max_seconds]
Power-of-two scaling of either operand is exact inside the admitted window (rt_selftest_range probes its edges), so one
exponent pair covers every pair of the window; the random self test (rt_selftest) keeps covering signs and exponents."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtb200

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/div_exhaustive.json"
ex_x, ex_d, signs = (int(v) for v in sys.argv[2:5]) if len(sys.argv) > 4 else (0, 0, 0)
budget = float(sys.argv[5]) if len(sys.argv) > 5 else 600.0
ctx = rtb200.Context(0)
CH = 32768
total_md = 1 << 23
bad = done = 0
t0 = time.time()
for md in range(0, total_md, CH):
    bad += ctx.selftest_exhaustive(md, min(CH, total_md - md), ex_x, ex_d, signs)
    done += min(CH, total_md - md)
    if time.time() - t0 > budget:
        break
dt = time.time() - t0
res = {"what": "div_hoisted(x, d, div_prepare(d)) and its packed FMUL2 + 2 FFMA2 form (both halves) vs x / d, bitwise",
       "exponents": {"x": ex_x, "d": ex_d}, "signs": signs, "divisor_mantissas_done": done, "of": total_md,
       "quotients_checked": done * (1 << 23), "comparisons": 3 * done * (1 << 23), "mismatches": bad, "seconds": dt,
       "complete": done == total_md}
print(json.dumps(res))
os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
json.dump(res, open(out, "w"), indent=1)
