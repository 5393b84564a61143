import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neurst_b200 import lib
M, N, K = [int(x) for x in sys.argv[1:4]]
A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
B = torch.randn(N, K, device="cuda").to(torch.bfloat16)
C = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    lib.gemm(A, B, C)
torch.cuda.synchronize()
