"""A/B aid: BA solves/s of the C2 window for whatever library / env the process was started with."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
r = bench.ba_bench(0, 1965.0, cpu_seconds=3.0)
print(sys.argv[1], "solves/s %.1f  ms %.4f  device ms %.4f  step launch us %.1f  4 windows %.0f  cpu %.1f  speedup %.2f" % (
    r["value"], r["ms_per_solve"], r["device_ms_per_solve"], r["roofline"]["launch_us"], r["concurrent_streams"]["value"], r["cpu_baseline"]["value"], r["speedup_vs_cpu"]))
