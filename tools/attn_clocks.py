"""SM clock, power and throttle reasons while ONE attention kernel runs back to back for a few seconds (is attention power-limited?).
    python tools/attn_clocks.py [fwd|bwd]"""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from aozora_sdxl_training_b200 import ops
which = sys.argv[1] if len(sys.argv) > 1 else "fwd"
B, H, T = 4, 10, 4096
q, k, v, do = [torch.randn(B, T, H, 64, device="cuda").to(torch.bfloat16) for _ in range(4)]
o, lse = ops.attn_fwd(q, k, v, 0.125)
fn = (lambda: ops.attn_fwd(q, k, v, 0.125)) if which == "fwd" else (lambda: ops.attn_bwd(q, k, v, o, do, lse, 0.125))
for _ in range(5):
    fn()
torch.cuda.synchronize()
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,temperature.gpu",
                      "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
t0 = time.time()
n = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < 4.0:
    for _ in range(50):
        fn()
    n += 50
    torch.cuda.synchronize()
e1.record()
torch.cuda.synchronize()
p.terminate()
out = p.communicate()[0].strip().splitlines()
ms = e0.elapsed_time(e1) / n
f = 4.0 * B * H * T * T * 64 * (1.0 if which == "fwd" else 2.5)
print(f"{which}: {ms * 1e3:.1f} us per call over {n} calls, {f / ms / 1e9:.1f} TFLOP/s sustained")
print("clock MHz, power W, sw_power_cap, hw_slowdown, temp C (100 ms samples):")
for l in out[:: max(1, len(out) // 12)]:
    print("  ", l)
