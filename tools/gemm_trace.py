import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from neurst_b200 import lib
L = lib.load()
M, N, K = [int(x) for x in sys.argv[1:4]]
A = torch.randn(M, K, device="cuda").to(torch.bfloat16); B = torch.randn(N, K, device="cuda").to(torch.bfloat16)
C = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(3): lib.gemm(A, B, C)
torch.cuda.synchronize()
tr = torch.zeros(2 * 296 * 8 * 8, dtype=torch.int64, device="cuda")
os.environ["B200ST_DEBUG_TRACE_PTR"] = str(tr.data_ptr())
lib.gemm(A, B, C); torch.cuda.synchronize()
t = tr[:296 * 64].view(296, 8, 8).cpu()
t2 = tr[296 * 64:].view(296, 8, 8).cpu()
for cta in (0, 1, 77):
    base = int(t[cta, 0, 0])
    print("CTA", cta)
    for tile in range(5):
        r = [int(x) - base if int(x) else -1 for x in t[cta, tile]]
        print("  tile %d: tma_first=%6d mma_enter=%6d acc_free=%6d first_full=%6d mma_done_issue=%6d epi_enter=%6d epi_tfull=%6d epi_done=%6d" % ((tile,) + tuple(r)))

for cta in (0, 77):
    for tile in range(3):
        r = [int(x) for x in t2[cta, tile]]
        print("CTA %d tile %d chunk0: ld_wait=%d issue_next=%d process=%d | chunk1: wait=%d process=%d" % (cta, tile, r[1]-r[0], r[2]-r[1], r[3]-r[2], r[4]-r[3], r[5]-r[4]))
