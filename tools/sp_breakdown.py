"""Where a (sequence-parallel) rank's step goes: per-kernel device time of the steady-state CUDA-graph replay, grouped
into GEMM / attention / small kernels / collectives, plus the idle time between kernels (span - sum of kernel time).

    python tools/sp_breakdown.py                                            # 1 GPU
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/sp_breakdown.py

Rank 0 profiles 5 replays with torch.profiler (CUPTI kernel records) and writes gpurun_out/sp_breakdown_N<world>.json.
"""
import argparse
import json
import os
import sys
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch

from alphagenome_b200.init_weights import apply_synth_weights
from alphagenome_b200.model import AlphaGenome, set_update_running_var
from alphagenome_b200.trunk import engine_for


def category(name: str) -> str:
    n = name.lower()
    if "nccl" in n:
        return "collective"
    if "gemm_conv_kernel" in n:
        return "gemm"
    if "attention_kernel" in n:
        return "attention"
    if n.startswith("void agb::") or "agb::" in n:
        return "small (libagb200)"
    if "memcpy" in n or "memset" in n:
        return "memcpy/memset"
    return "torch glue (cat / copy / index)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--length", type=int, default=1_048_576)
    ap.add_argument("--replays", type=int, default=5)
    ap.add_argument("--out", default=None, help="output file name under gpurun_out/ (default sp_breakdown_N<world>.json)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    model = AlphaGenome().to(dev).eval()
    set_update_running_var(model, False)
    apply_synth_weights(model, seed=0)
    eng = engine_for(model.transformer_unet, model)
    if world > 1:
        from alphagenome_b200.parallel import SeqParallel

        eng.sp = SeqParallel()
    model.enable_cuda_graphs(True)
    seq = torch.from_numpy(np.random.default_rng(42).integers(0, 4, size=(1, args.length)).astype(np.int64)).to(dev)
    org = torch.zeros(1, dtype=torch.long, device=dev)
    for _ in range(4):
        model(seq, org)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.replays):
        model(seq, org)
    e1.record()
    torch.cuda.synchronize()
    ms_plain = e0.elapsed_time(e1) / args.replays
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(args.replays):
            model(seq, org)
        torch.cuda.synchronize()
    if rank == 0:
        per_cat = defaultdict(lambda: [0.0, 0])
        per_kernel = defaultdict(lambda: [0.0, 0])
        t_min, t_max = None, None
        for ev in prof.events():
            if ev.device_type != torch.autograd.DeviceType.CUDA:
                continue
            dur = ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
            if dur is None or dur <= 0:
                continue
            c = category(ev.name)
            per_cat[c][0] += dur
            per_cat[c][1] += 1
            short = ev.name.split("(")[0][:90]
            per_kernel[short][0] += dur
            per_kernel[short][1] += 1
            s, e = ev.time_range.start, ev.time_range.end
            t_min = s if t_min is None else min(t_min, s)
            t_max = e if t_max is None else max(t_max, e)
        R = args.replays
        span_us = (t_max - t_min) / R if t_min is not None else None
        busy = sum(v[0] for v in per_cat.values()) / R
        out = dict(world=world, length_bp=args.length, ms_per_step_events=ms_plain, span_us_per_step_profiled=span_us,
                   kernel_us_per_step=busy, idle_us_per_step=(span_us - busy) if span_us else None,
                   categories={k: dict(us_per_step=v[0] / R, launches_per_step=v[1] / R) for k, v in sorted(per_cat.items(), key=lambda kv: -kv[1][0])},
                   top_kernels={k: dict(us_per_step=v[0] / R, launches_per_step=v[1] / R) for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1][0])[:25]})
        (ROOT / "gpurun_out").mkdir(exist_ok=True)
        (ROOT / "gpurun_out" / (args.out or f"sp_breakdown_N{world}.json")).write_text(json.dumps(out, indent=1))
        print("SP_BREAKDOWN " + json.dumps({k: out[k] for k in ("world", "ms_per_step_events", "span_us_per_step_profiled", "kernel_us_per_step", "idle_us_per_step", "categories")}))
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
