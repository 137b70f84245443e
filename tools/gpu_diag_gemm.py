"""GPU bring-up diagnostic: run each GEMM case in its own subprocess (a device trap must not poison the
rest), print one JSON line per case, then a short conv throughput probe.  Output: gpurun_out/diag_gemm.jsonl"""
import json
import os
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
OUT = ROOT / "gpurun_out"
OUT.mkdir(exist_ok=True)


def child(idx: int):
    import torch
    from alphagenome_b200 import _lib
    from tests.gemm_cases import CASES, run_case

    case = CASES[idx]
    try:
        r = run_case(**case)
    except Exception as e:  # noqa: BLE001
        r = dict(name=case["name"], ok=False, error=repr(e)[:500], watchdog=[hex(x) for x in _lib.watchdog_status(0)])
    print("RESULT " + json.dumps(r), flush=True)


def perf():
    import torch
    from alphagenome_b200 import _lib, ops

    dev = torch.device("cuda")
    res = []
    for (M, K, N, taps) in [(131072, 768, 768, 5), (65536, 1024, 1024, 5), (16384, 1536, 1536, 5), (131072, 768, 768, 1), (8192, 1536, 3072, 1)]:
        A = torch.randn(1, M, K, device=dev).bfloat16()
        W = (torch.randn(taps, N, K, device=dev) * 0.02).bfloat16()
        out = torch.empty(1, M, N, device=dev)
        for _ in range(3):
            ops.gemm(A, W, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 10
        e0.record()
        for _ in range(iters):
            ops.gemm(A, W, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        tf = 2.0 * M * K * N * taps / ms / 1e9
        res.append(dict(M=M, K=K, N=N, taps=taps, ms=ms, tflops=tf))
        print("PERF " + json.dumps(res[-1]), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child(int(sys.argv[2]))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "perf":
        perf()
        sys.exit(0)
    from tests.gemm_cases import CASES

    lines = []
    for i, c in enumerate(CASES):
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, __file__, "child", str(i)], capture_output=True, text=True, timeout=180)
            got = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
            if got:
                rec = json.loads(got[-1][7:])
            else:
                rec = dict(name=c["name"], ok=False, error="no result", rc=p.returncode, stderr=p.stderr[-800:])
        except subprocess.TimeoutExpired:
            rec = dict(name=c["name"], ok=False, error="timeout")
        rec["secs"] = round(time.time() - t0, 1)
        print(json.dumps(rec), flush=True)
        lines.append(rec)
    try:
        p = subprocess.run([sys.executable, __file__, "perf"], capture_output=True, text=True, timeout=300)
        for l in p.stdout.splitlines():
            if l.startswith("PERF "):
                print(l, flush=True)
                lines.append(json.loads(l[5:]))
        if p.returncode != 0:
            print("perf rc", p.returncode, p.stderr[-800:])
    except subprocess.TimeoutExpired:
        print("perf timeout")
    (OUT / "diag_gemm.jsonl").write_text("\n".join(json.dumps(l) for l in lines) + "\n")
