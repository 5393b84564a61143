"""Two eager training steps at the bench configuration (ncu target: no graph, no profiler)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from neurst_b200.trainer import build_speech_transformer_trainer, synthetic_batch
tr, _ = build_speech_transformer_trainer("speech_transformer_s", 8192, precision=os.environ.get("B200ST_TOOL_DTYPE", "fp16"), label_smoothing=0.1)
b = synthetic_batch(32, 1000, 88, 8192, device="cuda")
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    loss = tr.train_step(b, seed=i + 1)
torch.cuda.synchronize()
print("loss", float(loss))
