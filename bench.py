#!/usr/bin/env python
"""Benchmark of the B200 AlphaGenome trunk forward (BASELINE.json metric: trunk forward bp/s at 1 Mbp context).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg4|cfg4h|cfg3|cfg2] [--dp] [--impl reference]

A "step" is one forward  AlphaGenome(dna, organism_index)  over one batch of synthetic DNA tokens (uniform A/C/G/T,
np.random.default_rng(42)) with synthetic weights of the published architecture.  Workloads (BASELINE.json configs):

    cfg4   B=1 x 1,048,576 bp, trunk + output embeddings -> Embeds             (headline; the default)
    cfg4h  cfg4 + publication heads of human and mouse                         (config 4 as written)
    cfg3   B=1 x 131,072 bp + publication heads                                (config 3; --dp: one replica per GPU)
    cfg2   B=2 x 8,192 bp, trunk only                                          (config 2)

With N > 1 the default shards ONE context along the sequence over the N ranks (strong scaling); `--dp` runs N
independent replicas instead (weak scaling, no data-path collective).  Prints ONE JSON line (rank 0); the default run
also carries `secondary` (the other configs, measured in the same process) and `parity` (the B200 outputs against the
CPU oracle forward that `cpu_baseline` times).  See DESIGN.md "Measurement" for what each key means.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "trunk_forward_bp_per_s"
UNIT = "bp/s"
FULL_L = 1_048_576
CPU_SAMPLE_L = int(os.environ.get("AGB_BENCH_CPU_SAMPLE_BP", 131_072))   # the CPU arms time a bounded window of the workload
N_SITES = 64   # splice donor / acceptor sites per sample in the heads workloads

WORKLOADS = {
    "cfg4": dict(B=1, L=FULL_L, heads=False),
    "cfg4h": dict(B=1, L=FULL_L, heads=True),
    "cfg3": dict(B=1, L=131_072, heads=True),
    "cfg2": dict(B=2, L=8192, heads=False),
}


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index = index
        self.rows = []      # (host arrival time, line)
        self.proc = None
        self.t0 = self.t1 = None

    def start(self, wait_s: float = 3.0):
        """Launch nvidia-smi (its start-up alone can outlast a 0.3 s timed region) and wait for its first row; the rows
        that count are those arriving between mark_begin() and mark_end()."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            t_end = time.time() + wait_s
            while not self.rows and time.time() < t_end:
                time.sleep(0.02)
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        if self.t1 is None:
            self.mark_end()
        time.sleep(0.12)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        lo = self.t0 if self.t0 is not None else 0.0
        # a row describes the ~50 ms before it was printed: accept rows printed up to one period after the region ended
        rows = [r for t, r in self.rows if lo <= t <= self.t1 + 0.06]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=sorted(reasons), samples=len(sm))


def synth_tokens(B, L):
    return torch.from_numpy(np.random.default_rng(42).integers(0, 4, size=(B, L)).astype(np.int64))


def synth_sites(B, L):
    rs = np.random.default_rng(7)
    return (torch.from_numpy(np.sort(rs.choice(L, size=(B, N_SITES)), axis=1).astype(np.int64)),
            torch.from_numpy(np.sort(rs.choice(L, size=(B, N_SITES)), axis=1).astype(np.int64)))


# ----------------------------------------------------------------------------------------------------- CPU arms
def cpu_oracle_forward(L, threads, repeats=1, warmup=0):
    """The pinned CPU oracle (a port of the reference forward, see oracle/trunk_oracle.py) on the host cores: one
    B=1 x L bp trunk forward with the same synthetic weights (seed 0) and tokens as the B200 arm.
    Returns (bp/s, seconds, outputs of the last forward)."""
    from alphagenome_b200.init_weights import synth_state_dict
    from alphagenome_b200.model import AlphaGenome
    from oracle import trunk_oracle as O

    torch.set_num_threads(threads)
    m = AlphaGenome()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items() if v.is_floating_point()}
    sd = {k: v for k, v in m.state_dict().items()}
    sd.update(synth_state_dict(shapes, seed=0))
    del m
    seq = synth_tokens(1, L)
    org = torch.zeros(1, dtype=torch.long)
    times, out = [], None
    for i in range(warmup + repeats):
        t0 = time.perf_counter()
        out = O.alphagenome_embeds(sd, seq, org)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return L / float(np.median(times)), float(np.median(times)), out


def cpu_sample_note(L):
    return (f"B=1 x {CPU_SAMPLE_L} bp window" + (f" ({CPU_SAMPLE_L / L:.3g} of the {L} bp workload's context; bp/s on the CPU falls "
            "slowly with length, so the window flatters the CPU)" if CPU_SAMPLE_L != L else "") + ", fp32")


def run_reference_arm(args):
    """`--impl reference`: the reference's CPU forward for the same metric.  The reference is pure PyTorch whose imports
    (einx, torch_einops_utils, PoPE_pytorch) are not installable offline, so it cannot travel to the GPU box; the arm
    times the oracle port that is pinned to it, with every host thread, each step a bounded
    sample (one CPU_SAMPLE_L window) of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.config]
    cores = os.cpu_count() or 1
    bps, secs, _ = cpu_oracle_forward(CPU_SAMPLE_L, cores, repeats=args.steps, warmup=min(args.warmup, 1))
    sample = cpu_sample_note(w["L"]) + f", {cores} threads, trunk + output embeddings"
    line = dict(metric=METRIC, value=bps, unit=UNIT, impl="reference", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=secs * 1e3, higher_is_better=True, scaling=scaling_label(args), vs_baseline=None, dtype="f32", data="synthetic",
                config=workload_config(args.config, args.gpus, args.dp),
                cpu_baseline=dict(value=bps, unit=UNIT, cores=cores, kind="port", sample=sample),
                e2e=dict(value=bps, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


def scaling_label(args):
    # the default shards ONE fixed context over the ranks (total work fixed); --dp gives every rank its own replica
    return "weak" if args.dp else "strong"


def workload_config(name, gpus, dp=False):
    w = WORKLOADS[name]
    B, L = w["B"], w["L"]
    if gpus == 1:
        par = "single GPU"
    elif dp:
        par = f"dp{gpus}: {gpus} independent replicas, one per GPU (batch data parallel; no data-path collective)"
    else:
        par = (f"sp{gpus}: ONE {L} bp context sharded along the sequence over {gpus} GPUs (2048 bp recomputed conv halos, "
               f"K/V + pooled-row all-gathers and one pair all-to-all over NCCL/NVLink)")
    return dict(workload="AlphaGenome() forward (conv encoder + 9-layer tower with pair blocks + conv decoder + output "
                         f"embeddings), B={B} x {L} bp, synthetic weights (default 7-stage dims 768..1536), "
                         + (f"+ publication heads of human and mouse (add_heads(**publication_heads_config[...]), {N_SITES} splice sites)"
                            if w["heads"] else "no heads"),
                name=name, length_bp=L, global_batch=B * (gpus if dp else 1), parallelism=par, cpu_sample_bp=CPU_SAMPLE_L,
                l2="per-step weights (0.8 GB bf16) + activations exceed the 126 MB L2 at every config; no explicit flush needed")


# ----------------------------------------------------------------------------------------------------- GPU arm
class Workload:
    """One model + resident synthetic inputs; step() is one forward through the public API."""

    def __init__(self, name, dev, sp=None, graphs=True):
        from alphagenome_b200.init_weights import apply_synth_weights
        from alphagenome_b200.model import AlphaGenome, publication_heads_config, set_update_running_var
        from alphagenome_b200.trunk import engine_for

        w = WORKLOADS[name]
        self.name, self.B, self.L, self.heads = name, w["B"], w["L"], w["heads"]
        model = AlphaGenome()
        if self.heads:
            for organism in ("human", "mouse"):
                model.add_heads(**publication_heads_config[organism])
        self.model = model.to(dev).eval()
        set_update_running_var(self.model, False)   # inference: frozen BatchRMSNorm statistics, as the reference's inference tests run it
        apply_synth_weights(self.model, seed=0)
        self.eng = engine_for(self.model.transformer_unet, self.model)
        self.eng.sp = sp
        self.model.enable_cuda_graphs(graphs)
        self.tok_host = synth_tokens(self.B, self.L).pin_memory()
        self.org_host = (torch.arange(self.B, dtype=torch.long) % 2).pin_memory()
        self.tok = self.tok_host.to(dev)
        self.org = self.org_host.to(dev)
        self.kw = {}
        if self.heads:
            d, a = synth_sites(self.B, self.L)
            self.kw = dict(splice_donor_idx=d.to(dev), splice_acceptor_idx=a.to(dev))

    def step(self):
        return self.model(self.tok, self.org, **self.kw)

    def pipeline(self):
        from alphagenome_b200.pipeline import HostPipeline

        return HostPipeline(self.model, forward_kwargs=self.kw)   # reads back EVERYTHING the call returns


def barrier(world):
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
    torch.cuda.synchronize()


def time_steps(fn, steps, world):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(world)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    barrier(world)
    return e0.elapsed_time(e1) / steps


def max_over_ranks(vals, dev, world):
    t = torch.tensor(vals, device=dev, dtype=torch.float64)
    if world > 1:
        import torch.distributed as dist

        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def time_e2e(wl, steps, world):
    """Through the public API with HOST buffers: per step H2D of the pinned tokens, forward, D2H of everything the call
    returns into pinned host buffers (double-buffered behind the next forward)."""
    pipe = wl.pipeline()
    pipe.run((wl.tok_host, wl.org_host) for _ in range(2))
    pipe.h2d_bytes = pipe.d2h_bytes = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(world)
    e0.record()
    pipe.run((wl.tok_host, wl.org_host) for _ in range(steps))
    e1.record()
    barrier(world)
    ms = e0.elapsed_time(e1) / steps
    return ms, pipe.h2d_bytes // steps, pipe.d2h_bytes // steps, pipe.names


def secondary_workload(name, dev, world, steps, warmup, with_e2e=False):
    """Another BASELINE config measured in the same process (this rank's own replica; max over ranks)."""
    torch.cuda.empty_cache()
    wl = Workload(name, dev)
    for _ in range(warmup):
        wl.step()
    ms = time_steps(wl.step, steps, world)
    out = dict(ms_per_step=ms)
    if with_e2e:
        ms_e, h2d, d2h, _ = time_e2e(wl, steps, world)
        out.update(e2e_ms_per_step=ms_e, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h)
    vals = max_over_ranks([out["ms_per_step"], out.get("e2e_ms_per_step", 0.0)], dev, world)
    out["ms_per_step"] = vals[0]
    out["bp_per_s"] = world * wl.B * wl.L / (vals[0] * 1e-3)
    if with_e2e:
        out["e2e_ms_per_step"] = vals[1]
        out["e2e_bp_per_s"] = world * wl.B * wl.L / (vals[1] * 1e-3)
    out["replicas"] = world
    del wl
    torch.cuda.empty_cache()
    return out


def parity_vs_oracle(dev, ref):
    """The B200 forward on the CPU arm's sample (same weights, same tokens) against the oracle outputs `ref`."""
    from alphagenome_b200.init_weights import apply_synth_weights
    from alphagenome_b200.model import AlphaGenome, set_update_running_var

    model = AlphaGenome().to(dev).eval()
    set_update_running_var(model, False)
    apply_synth_weights(model, seed=0)
    emb = model(synth_tokens(1, CPU_SAMPLE_L).to(dev), torch.zeros(1, dtype=torch.long, device=dev))
    torch.cuda.synchronize()
    rep = {}
    for name, got in zip(("embeds_1bp", "embeds_128bp", "embeds_pair"), emb):
        g, w = got.float().cpu(), ref[name]
        a, b = g.double().flatten(), w.double().flatten()
        a, b = a - a.mean(), b - b.mean()
        rep[name] = dict(max_rel=((g - w).abs().max() / w.abs().max()).item(), pearson=float((a @ b) / (a.norm() * b.norm())))
    return dict(against="CPU oracle forward of the cpu_baseline leg", length_bp=CPU_SAMPLE_L, tolerance="max|d|/max|ref| <= 1e-2, Pearson >= 0.999",
                ok=all(v["max_rel"] <= 1e-2 and v["pearson"] >= 0.999 for v in rep.values()), **rep)


def measured_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant GEMM shape, from the committed ncu
    --set full capture (profiles/gemm_traffic.json is written by tools/ncu_summary.py from the .ncu-rep)."""
    p = ROOT / "profiles" / "gemm_traffic.json"
    if not p.exists():
        return None, "no ncu capture committed"
    d = json.loads(p.read_text())
    return d.get("dram_bytes_per_launch"), d.get("note", "")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--heads", action="store_true", help="alias of --config cfg4h (trunk + publication heads at 1 Mbp)")
    ap.add_argument("--dp", action="store_true", help="N > 1: independent replicas (batch data parallel) instead of sharding one context")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the other BASELINE configs the default run also measures")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
    args = ap.parse_args()
    if args.heads:
        args.config = "cfg4h"
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    from alphagenome_b200 import _lib, ops

    _lib.check(_lib.load().ag_check_device(local_rank))
    sp = None
    if world > 1 and not args.dp:
        from alphagenome_b200.parallel import SeqParallel

        sp = SeqParallel()  # strong scaling: the same context, sharded along the sequence
    wl = Workload(args.config, dev, sp=sp, graphs=not args.no_graph)
    B, L = wl.B, wl.L
    units = B * L * (world if args.dp else 1)     # bp the whole job processes per step

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        wl.step()
    barrier(world)

    # --- timed: device-resident inputs ---
    sampler.mark_begin()
    ms = time_steps(wl.step, args.steps, world)
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None

    # --- timed: end to end through the public API with host buffers ---
    ms_e2e, h2d, d2h, e2e_names = time_e2e(wl, args.steps, world)

    # --- dominant-kernel timing: K more steps launched eagerly with a CUDA-event pair around every ag_gemm_fwd
    #     (events cannot be recorded inside a graph replay); also counts our kernel launches per step ---
    wl.model.enable_cuda_graphs(False)
    wl.step()
    barrier(world)
    ops.PROFILE_GEMM = []
    launches0 = _lib.launch_count()
    ms_eager = time_steps(wl.step, args.steps, world)
    launches = _lib.launch_count() - launches0
    prof = ops.PROFILE_GEMM
    ops.PROFILE_GEMM = None
    times = [a.elapsed_time(b) for a, b, _, _ in prof]
    gemm_ms = sum(times)
    gemm_flops = sum(f for _, _, f, _ in prof)
    n_gemm = len(prof)
    # the same launches split by arithmetic intensity against the machine's ridge point (sustained bf16 peak / HBM peak):
    # convs and wide linears are tensor-bound, the K <= 128 pair linears / w1 convs / embedding GEMMs are HBM-bound
    pk = peaks()
    ridge = pk["tf_sustained"] * 1e12 / (pk["hbm"] * 1e9)
    t_ms = sum(t for t, (_, _, f, nb) in zip(times, prof) if f / nb >= ridge)
    t_fl = sum(f for _, _, f, nb in prof if f / nb >= ridge)
    h_ms = sum(t for t, (_, _, f, nb) in zip(times, prof) if f / nb < ridge)
    h_by = sum(nb for _, _, f, nb in prof if f / nb < ridge)
    n_tensor = sum(1 for _, _, f, nb in prof if f / nb >= ridge)
    wl.model.enable_cuda_graphs(not args.no_graph)

    ms, ms_e2e, ms_eager = max_over_ranks([ms, ms_e2e, ms_eager], dev, world)

    # --- the other BASELINE configs, in the same process (every rank runs its own replica: data parallel) ---
    secondary = None
    if not args.no_secondary and args.config == "cfg4":
        k = min(args.steps, 10)
        secondary = {}
        if world == 1:
            del wl
            torch.cuda.empty_cache()
            secondary["cfg4h_trunk_plus_heads_1mbp"] = secondary_workload("cfg4h", dev, world, k, 3)
            secondary["cfg2_B2_8192_trunk"] = secondary_workload("cfg2", dev, world, max(k, 20), 5, with_e2e=True)
        secondary["cfg3_131k_plus_heads" + (f"_dp{world}" if world > 1 else "")] = secondary_workload("cfg3", dev, world, max(k, 10), 3, with_e2e=True)
        hm = secondary.get("cfg4h_trunk_plus_heads_1mbp")
        secondary["summary"] = dict(heads_ms=(hm["ms_per_step"] - ms) if hm else None,
                                    cfg2_bp_s=secondary.get("cfg2_B2_8192_trunk", {}).get("bp_per_s"),
                                    cfg3_bp_s=next(v["bp_per_s"] for k_, v in secondary.items() if k_.startswith("cfg3")))

    if rank != 0:
        finish(world)
        return
    value = units / (ms * 1e-3)
    e2e = units / (ms_e2e * 1e-3)
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
    traffic, traffic_note = measured_traffic() if (args.config == "cfg4" and world == 1) else (None, "")
    graphed = not args.no_graph
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=ms,
                higher_is_better=True, scaling=scaling_label(args), vs_baseline=None,
                dtype="bf16 operands / fp32 accumulate + fp32 residual stream",
                data="synthetic", config=workload_config(args.config, world, args.dp),
                e2e=dict(value=e2e, unit=UNIT, ms_per_step=ms_e2e, h2d_bytes_per_step=h2d * world, d2h_bytes_per_step=d2h * world,
                         d2h_gb_per_s_per_gpu=d2h / (ms_e2e * 1e-3) / 1e9, outputs=e2e_names,
                         how="alphagenome_b200.pipeline.HostPipeline around model(dna, organism_index): per step H2D of the pinned host "
                             "tokens, forward, D2H of EVERY tensor the call returns (fp32, as the reference returns them) into pinned "
                             "host buffers, double-buffered behind the next forward; PCIe-bound when the read-back outlasts the forward"),
                gpu_launches=int(launches), launch_mode="cuda-graph replay" if graphed else "eager",
                clocks=clocks,
                roofline=dict(kernel="gemm_conv_kernel (tcgen05 implicit-GEMM conv/linear), rank 0", bound="tensor", achieved=achieved,
                              peak=pk["tf_sustained"], unit="TFLOP/s", frac=(achieved / pk["tf_sustained"]) if achieved else None,
                              traffic=traffic, traffic_note=traffic_note,
                              peak_source=pk["src"] + " (sustained cuBLAS bf16; kernel timed inside a long step)",
                              launches_per_step=n_gemm // max(args.steps, 1), share_of_step=gemm_ms / (ms_eager * args.steps),
                              algorithmic_tflop_per_step=gemm_flops / max(args.steps, 1) / 1e12,
                              by_bound=dict(
                                  ridge_flop_per_byte=ridge,
                                  tensor_bound=dict(launches_per_step=n_tensor // max(args.steps, 1), share_of_gemm_time=t_ms / gemm_ms if gemm_ms else None,
                                                    achieved_tflops=t_fl / (t_ms * 1e-3) / 1e12 if t_ms else None,
                                                    frac=t_fl / (t_ms * 1e-3) / 1e12 / pk["tf_sustained"] if t_ms else None),
                                  hbm_bound=dict(launches_per_step=(n_gemm - n_tensor) // max(args.steps, 1), share_of_gemm_time=h_ms / gemm_ms if gemm_ms else None,
                                                 achieved_gbs=h_by / (h_ms * 1e-3) / 1e9 if h_ms else None,
                                                 frac=h_by / (h_ms * 1e-3) / 1e9 / pk["hbm"] if h_ms else None,
                                                 note="algorithmic bytes (each operand / output once); eager launches, so the many "
                                                      "sub-100 us pair linears also carry their launch gaps")),
                              measured="CUDA events around every launch in %d eagerly launched steps (%.2f ms/step) right after the timed region" % (args.steps, ms_eager)))
    if secondary is not None:
        line["secondary"] = secondary
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        bps, secs, ref = cpu_oracle_forward(CPU_SAMPLE_L, cores)
        line["cpu_baseline"] = dict(value=bps, unit=UNIT, cores=cores, kind="port", seconds=secs,
                                    sample="one trunk forward of " + cpu_sample_note(L) + " through the pinned CPU oracle")
        line["parity"] = parity_vs_oracle(dev, ref)
    print(json.dumps(line), flush=True)
    finish(world)


def finish(world):
    """Leave a multi-rank run without a collective teardown.  Every rank synchronises its device, meets the others at a
    last barrier and exits the process directly: destroy_process_group() / interpreter finalisation with captured NCCL
    kernels still referenced by CUDA graphs hung the 8-GPU run after the result line had been printed."""
    if world <= 1:
        return
    import torch.distributed as dist

    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
