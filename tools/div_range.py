import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtb200
ctx = rtb200.Context(0)
for ex in list(range(-149, -90, 3)) + list(range(-90, 60, 10)) + [60, 80, 100, 105, 110, 120, 127]:
    print(ex, ctx.selftest_range(1 << 24, ex), flush=True)
