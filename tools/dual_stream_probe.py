"""Hypothesis probe: two half-batch training steps replayed concurrently on two streams vs one full-batch step."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from neurst_b200.trainer import build_speech_transformer_trainer, synthetic_batch

def timed(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

nsplit = int(sys.argv[1]) if len(sys.argv) > 1 else 2
full, _ = build_speech_transformer_trainer("speech_transformer_s", 8192, precision="bf16", label_smoothing=0.1, use_cuda_graph=True)
bf = synthetic_batch(32, 1000, 88, 8192, device="cuda")
ctr = [0]
def step_full():
    ctr[0] += 1; full.train_step(bf, seed=ctr[0])
print("full B=32: %.3f ms" % timed(step_full))
del full
parts = [build_speech_transformer_trainer("speech_transformer_s", 8192, precision="bf16", label_smoothing=0.1, use_cuda_graph=True)[0] for _ in range(nsplit)]
bs = [synthetic_batch(32 // nsplit, 1000, 88, 8192, device="cuda", seed=7 + i) for i in range(nsplit)]
streams = [torch.cuda.Stream() for _ in range(nsplit)]
for t, b in zip(parts, bs): t.train_step(b, seed=1)      # capture
torch.cuda.synchronize()
def step_split():
    ctr[0] += 1
    cur = torch.cuda.current_stream()
    for t, b, s in zip(parts, bs, streams):
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            t.train_step(b, seed=ctr[0])
    for s in streams: cur.wait_stream(s)
print("%d x B=%d concurrent: %.3f ms" % (nsplit, 32 // nsplit, timed(step_split)))
def step_serial():
    ctr[0] += 1
    for t, b in zip(parts, bs): t.train_step(b, seed=ctr[0])
print("%d x B=%d serial: %.3f ms" % (nsplit, 32 // nsplit, timed(step_serial)))
