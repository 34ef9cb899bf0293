"""Host vs device time of bench.full_training_step (run on the GPU box): python tools/time_full_step.py"""
import os, sys, time, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
t0 = time.perf_counter()
r = bench.full_training_step(dev, 1, B=int(os.environ.get("B", "2")), steps=4, warmup=2)
print({k: r[k] for k in ("ms_per_step", "value", "peak_memory_gib", "host_enqueue_ms_per_step")}, "total wall", round(time.perf_counter() - t0, 2))
if os.environ.get("NOPROF"):
    sys.exit(0)
pr = cProfile.Profile()
pr.enable()
r = bench.full_training_step(dev, 1, B=int(os.environ.get("B", "2")), steps=6, warmup=2)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
