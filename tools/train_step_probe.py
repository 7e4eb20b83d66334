"""Runs only the extra keys of bench.py (PanoSwin-B, Mask R-CNN forward, training step) and prints the training-step entry."""
import os
import sys, json, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv=["bench.py"]
import bench, torch
dev=torch.device("cuda",0)
def timed(fn, steps):
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
args=types.SimpleNamespace()
out=bench.run_extras(args, dev, 1, 0, timed)
print(json.dumps(out["train_step"], indent=1))
