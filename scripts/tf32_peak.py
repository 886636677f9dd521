"""Practical tensor-core ceilings on this box (cuBLAS through torch.matmul, 8192^3, sustained 3 s each): single-pass TF32 and
bf16, with SM clock / power sampled under load.  Context for the split-TF32 kernels: they execute 3 TF32 MMAs per
algorithmic multiply-add, so their hardware rate is 3x the algorithmic TFLOP/s bench.py reports."""
import subprocess, threading, time, torch
dev = torch.device("cuda:0")
def sample(stop, out):
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                               capture_output=True, text=True, timeout=5).stdout.strip().split(",")
            out.append((float(r[0]), float(r[1])))
        except Exception:
            pass
        time.sleep(0.2)
def run(dtype, tf32, n=8192, secs=3.0):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device=dev, dtype=dtype); b = torch.randn(n, n, device=dev, dtype=dtype)
    for _ in range(5): a @ b
    torch.cuda.synchronize()
    stop, smp = threading.Event(), []
    th = threading.Thread(target=sample, args=(stop, smp)); th.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 0; t0 = time.time(); e0.record()
    while time.time() - t0 < secs:
        for _ in range(20): a @ b
        iters += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize(); stop.set(); th.join()
    ms = e0.elapsed_time(e1) / iters
    clk = sorted(s[0] for s in smp)[len(smp) // 2] if smp else 0; pw = max(s[1] for s in smp) if smp else 0
    print("%s%s %d^3: %.1f TFLOP/s sustained, sm clock median %.0f MHz, power max %.0f W" %
          (str(dtype).split(".")[-1], " (tf32 tensor cores)" if tf32 else "", n, 2 * n ** 3 / ms / 1e9, clk, pw))
run(torch.float32, True)
run(torch.bfloat16, False)
run(torch.float32, False)
