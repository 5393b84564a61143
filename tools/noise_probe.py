"""Dev tool: run-to-run noise of the gradient arena (same batch, same dropout seed, two eager backward passes)."""
import sys
sys.path.insert(0, ".")
import torch
from neurst_b200.trainer import build_speech_transformer_trainer, synthetic_batch

for prec in ("fp16", "bf16", "fp32"):
    tr, _ = build_speech_transformer_trainer("speech_transformer_s", vocab_size=96, precision=prec, label_smoothing=0.1, seed=5)
    rt = tr.rt
    batch = synthetic_batch(4, 160, 12, 96, seed=100, device="cuda")
    b = dict(batch); b.update(training=True, seed=77, want_logits=False)
    gs = []
    for _ in range(3):
        rt.ensure_grads().zero_()
        rt.run(b, backward=True)
        torch.cuda.synchronize()
        gs.append(rt.grads.clone())
    n = gs[0].norm()
    print(prec, "rel diff run1-run0 %.3e  run2-run0 %.3e" % (float((gs[1] - gs[0]).norm() / n), float((gs[2] - gs[0]).norm() / n)), flush=True)
    # per-tensor worst
    worst = sorted(((float((rt.view(k, gs[1]) - rt.view(k, gs[0])).norm() / (rt.view(k, gs[0]).norm() + 1e-30)), k) for k in rt.table), reverse=True)[:4]
    print("   worst tensors:", worst, flush=True)
