"""ncu target: the frames pipeline (cps_fit_frames_batched) on the bench's frame batch."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_extras as b
print(json.dumps(b.frames_bench(0))[:600])
