import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from neurst_b200 import lib
L = lib.load()
dev = "cuda"
def rnd(*s): return torch.randn(*s, device=dev).to(torch.bfloat16)
def run(tag, M, N, K, out_dtype=torch.bfloat16, a_mn=False, b_mn=False, **kw):
    A = rnd(K, M) if a_mn else rnd(M, K)
    B = rnd(K, N) if b_mn else rnd(N, K)
    Cc = torch.zeros(M, N, device=dev, dtype=out_dtype)
    ms = lib.gemm_bench(A, B, Cc, iters=50, a_mn=a_mn, b_mn=b_mn, **kw)
    print("%-40s %8.1f us %8.1f TF" % (tag, ms * 1e3, 2.0 * M * N * K / ms / 1e9), flush=True)
bias = torch.randn(2048, device=dev)
res = torch.randn(8000, 2048, device=dev)
msk = rnd(8000, 2048)
for M in (8000, 8192):
    run("ffn1 plain M=%d" % M, M, 2048, 256)
    run("ffn1 b_mn M=%d" % M, M, 2048, 256, b_mn=True)
run("ffn1 bias", 8000, 2048, 256, b_mn=True, bias=bias)
run("ffn1 bias relu", 8000, 2048, 256, b_mn=True, bias=bias, relu=True)
run("ffn1 bias relu drop(philox)", 8000, 2048, 256, b_mn=True, bias=bias, relu=True, dropout=(0.1, 1, 2))
run("ffn1 mask", 8000, 2048, 256, mask_src=msk)
run("ffn1 f32 out", 8000, 2048, 256, out_dtype=torch.float32)
run("ffn1 f32 out + residual", 8000, 2048, 256, out_dtype=torch.float32, residual=res)
bias2 = torch.randn(256, device=dev); res2 = torch.randn(8000, 256, device=dev)
run("ffn2 plain bf16", 8000, 256, 2048)
run("ffn2 f32 bias res", 8000, 256, 2048, out_dtype=torch.float32, bias=bias2, residual=res2, b_mn=True)
run("out-proj f32 bias res", 8000, 256, 256, out_dtype=torch.float32, bias=bias2, residual=res2, b_mn=True)
run("out-proj plain bf16", 8000, 256, 256)
run("qkv bias", 8000, 768, 256, b_mn=True, bias=torch.randn(768, device=dev))
for sk in (1, 2, 4, 8, 16):
    run("wgrad 256x2048xK8000 splitk=%d" % sk, 256, 2048, 8000, out_dtype=torch.float32, a_mn=True, b_mn=True, accumulate=True, splitk=sk)
for sk in (1, 4, 8, 13, 26):
    run("wgrad 256x768xK8000 splitk=%d" % sk, 256, 768, 8000, out_dtype=torch.float32, a_mn=True, b_mn=True, accumulate=True, splitk=sk)
