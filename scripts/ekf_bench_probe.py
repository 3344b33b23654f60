"""EKF update timings at the BASELINE sizes (device events of the handle): M = 35 at 1k / 5k / 20k point landmarks,
M = 65 (4 curves + 10 points) at 1k / 5k.  Prints one JSON object."""
import json
import sys

sys.path.insert(0, ".")
import bench_extras as bx

out = {}
for name, L, npm in (("1k_M35", 1000, 0), ("5k_M35", 5000, 0), ("20k_M35", 20000, 0), ("1k_M65", 1000, 10),
                     ("5k_M65", 5000, 10), ("5k_M80", 5000, 15)):
    r = bx.ekf_update_bench(0, L, n_point_meas=npm)
    out[name] = {k: r[k] for k in ("N", "M", "update_ms", "cov_kernel_ms", "predict_ms", "launches_per_update")}
    out[name]["frac"] = r["roofline"]["frac"]
print(json.dumps(out))
