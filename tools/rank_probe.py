#!/usr/bin/env python
"""Per-rank diagnostics under torchrun: is a slow rank a slow GPU (clocks), the shard, or the peer stores?"""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import cutrace_b200 as ct
from cutrace_b200.distributed import TileShardedRenderer
import bench
import pynvml

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(lr)
clk, stop = [], threading.Event()
def sample():
    while not stop.is_set():
        clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)); stop.wait(0.02)
scene, wl = bench.load_workload(sys.argv[1] if len(sys.argv) > 1 else "bunny4k")
def timed(fn, n):
    torch.cuda.synchronize(); dist.barrier(); clk.clear()
    ms = []
    for _ in range(n):
        ms.append(fn())
    return float(np.median(ms)), (int(np.median(clk)) if clk else -1)
t = threading.Thread(target=sample, daemon=True); t.start()
with ct.Renderer(scene, device=lr) as r:
    r.render()
    full, c_full = timed(lambda: r.render()["render_ms"], 5)
with ct.Renderer(scene, device=lr, tile_rank=rank, tile_world=world) as r:
    r.render()
    alone, c_alone = timed(lambda: r.render()["render_ms"], 20)
    def step():
        ms = r.render()["render_ms"]; dist.barrier(); return ms
    sync, c_sync = timed(step, 20)
res = {}
for label, env in (("peer", {}), ("no-export", {"CUTRACE_DEBUG_SKIP_EXPORT": "1"}), ("no-export,local-colour", {"CUTRACE_DEBUG_SKIP_EXPORT": "1", "CUTRACE_DEBUG_LOCAL_COLOR": "1"}),
                   ("export,local-colour", {"CUTRACE_DEBUG_LOCAL_COLOR": "1"})):
    for k in ("CUTRACE_DEBUG_SKIP_EXPORT", "CUTRACE_DEBUG_LOCAL_COLOR"):
        os.environ.pop(k, None)
    os.environ.update(env)
    tsr = TileShardedRenderer(scene, rank=rank, world=world, device=lr, exchange="peer")
    tsr.render()
    def step2():
        ms = tsr.render()["render_ms"]; tsr.gather(); return ms
    res[label], _ = timed(step2, 20)
    tsr.close()
stop.set()
print(f"[rank {rank}] full-frame {full:.3f} ms @{c_full} MHz | shard no-sync {alone:.3f} | shard+barrier {sync:.3f} | " + " | ".join(f"{k} {v:.3f}" for k, v in res.items()), flush=True)
dist.destroy_process_group()
