"""Critical-path contribution of each kernel class to the cfg-2 training step: the step is re-captured and timed with the
launches of one class SKIPPED (B200ST_ABLATE bit mask, csrc/common.cuh) — the difference to the full step is what that
class costs on the critical path (side-stream work that is fully hidden shows ~0).  Results are garbage while a class is
skipped; this is a measurement aid only.

    python tools/ablate_step.py [--steps 10]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

CLASSES = [("full", 0), ("colsum", 1), ("wgrad_gemm", 2), ("ln_fwd", 4), ("ln_bwd", 8), ("attn_fwd", 16), ("attn_bwd", 32),
           ("mlp_fwd", 64), ("mlp_bwd", 128), ("other_gemm", 256), ("optimizer", 512), ("dropout_bits", 1024), ("conv1+im2col", 2048),
           ("colsum+wgrad (all side-stream GEMM work)", 3), ("ln_fwd+ln_bwd", 12), ("full_again", 0)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--dtype", default="fp16")
    args = ap.parse_args()
    from neurst_b200.trainer import build_speech_transformer_trainer, synthetic_batch
    tr, _ = build_speech_transformer_trainer("speech_transformer_s", 8192, precision=args.dtype, label_smoothing=0.1, seed=1234,
                                             use_cuda_graph=True)
    batches = [synthetic_batch(32, 1000, 88, 8192, seed=1234 + i, device="cuda") for i in range(4)]
    out = {}
    for name, mask in CLASSES:
        os.environ["B200ST_ABLATE"] = str(mask)
        tr._graphs.clear()
        for i in range(3):
            tr.train_step(batches[i % 4], seed=i + 1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            tr.train_step(batches[i % 4], seed=i + 1)
        e1.record()
        torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / args.steps
        print("%-44s mask %5d  %.3f ms/step   delta %+.3f" % (name, mask, out[name], out[name] - out["full"]), flush=True)
    os.environ["B200ST_ABLATE"] = "0"
    print(json.dumps(out))


if __name__ == "__main__":
    main()
