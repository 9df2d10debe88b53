"""Per-run wall times of the θ-sharded config 5 under torchrun (are timed runs repeatable?):
torchrun --nproc-per-node G tools/c5_runs_probe.py [runs]      SMI=1 also runs an `nvidia-smi -lms 100` poller per rank as bench.py does"""
import json, os, subprocess, sys
sys.path.insert(0, ".")
import torch
import torch.distributed as dist
import sequential_monte_carlo_b200 as smc
from sequential_monte_carlo_b200 import bench_smc2 as B

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
comm = smc.TorchComm() if world > 1 else None
ctx = smc.Context(local, 1998)
runs = int(sys.argv[1]) if len(sys.argv) > 1 else 6
poll, stop_flag, samples = None, [False], []
mode = os.environ.get("SMI", "0")
if mode == "1":            # a poller per rank at 100 ms (what bench.py did)
    poll = subprocess.Popen(["nvidia-smi", "-i", str(local), "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "100"],
                            stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
elif mode == "smi1000" and rank == 0:   # one poller for the box at 1 s
    poll = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "1000"],
                            stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
elif mode == "nvml" and rank == 0:      # NVML from a thread of rank 0, every 200 ms, all GPUs of the job
    import threading, time, pynvml
    pynvml.nvmlInit()
    hs = [pynvml.nvmlDeviceGetHandleByIndex(i) for i in range(world)]
    def loop():
        while not stop_flag[0]:
            for h in hs:
                samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h),
                                pynvml.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons")
                                else pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
            time.sleep(0.2)
    th = threading.Thread(target=loop, daemon=True)
    th.start()
barrier = (lambda: dist.barrier()) if world > 1 else None
out = B.run_config("c5", ctx, comm, rank, world, steps=runs, warmup=2, barrier=barrier)
if poll:
    poll.terminate()
stop_flag[0] = True
print(json.dumps({"rank": rank, "world": world, "smi": os.environ.get("SMI", "0"), "nvml_samples": len(samples), "walls_s": [round(w, 4) for w in out["walls_s"]], "spans_s": [round(w, 4) for w in out["device_spans_s"]]}), flush=True)
if world > 1:
    dist.destroy_process_group()
