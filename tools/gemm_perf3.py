import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from neurst_b200 import lib
L = lib.load()
def rnd(*s): return torch.randn(*s, device="cuda").to(torch.bfloat16)
for (M, N, K) in [(8192, 2048, 256), (8192, 768, 256), (8192, 256, 2048), (160000, 256, 2304), (8192, 8192, 8192)]:
    A, B = rnd(M, K), rnd(N, K)
    Cc = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    ms = lib.gemm_bench(A, B, Cc, iters=50)
    print("%s skip_store=%s: %.1f us %.1f TF" % ((M, N, K), os.environ.get("B200ST_DEBUG_SKIP_STORE"), ms * 1e3, 2.0 * M * N * K / ms / 1e9))
