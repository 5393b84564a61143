"""Dev tool: time the layer GEMM shapes under the epilogue debug modes (B200ST_EPI_MODE is read per launch).
Usage: python tools/gemm_epi_probe.py"""
import os, sys, json
sys.path.insert(0, ".")
import torch
from neurst_b200 import lib as L

def run(name, M, N, K, dt=torch.float16, out=torch.float16, a_mn=False, b_mn=True, bias=True, relu=False, acc=False, splitk=1, iters=50):
    A = (torch.randn((K, M) if a_mn else (M, K), device="cuda") * 0.1).to(dt)
    B = (torch.randn((K, N) if b_mn else (N, K), device="cuda") * 0.1).to(dt)
    C = torch.zeros(M, N, device="cuda", dtype=out)
    kw = dict(a_mn=a_mn, b_mn=b_mn, relu=relu, accumulate=acc, splitk=splitk)
    if bias:
        kw["bias"] = torch.randn(N, device="cuda")
    res = {}
    for mode in ("0", "1", "2", "3"):
        os.environ["B200ST_EPI_MODE"] = mode
        ms = L.gemm_bench(A, B, C, iters=iters, **kw)
        res[mode] = round(ms * 1e3, 2)
    os.environ["B200ST_EPI_MODE"] = "0"
    tf = 2.0 * M * N * K / (res["0"] * 1e-6) / 1e12
    print("%-28s M=%d N=%d K=%d  us: full %.1f | no-store %.1f | no-stage %.1f | no-tmem-ld %.1f   (%.0f TF/s)" %
          (name, M, N, K, res["0"], res["1"], res["2"], res["3"], tf), flush=True)
    return res

if __name__ == "__main__":
    L.load()
    out = {}
    out["ffn1"] = run("ffn1 fwd (bias,relu)", 8000, 2048, 256, relu=True)
    out["qkv"] = run("qkv fwd (bias)", 8000, 768, 256)
    out["out"] = run("out-proj (f32 out)", 8000, 256, 256, out=torch.float32)
    out["ffn2"] = run("ffn2 fwd (f32 out)", 8000, 256, 2048, out=torch.float32)
    out["ffn2_dgrad"] = run("ffn2 dgrad (f16 out)", 8000, 2048, 256, b_mn=False, bias=False)
    out["ffn1_dgrad"] = run("ffn1 dgrad (f32 out)", 8000, 256, 2048, b_mn=False, bias=False, out=torch.float32)
    out["wgrad_ffn"] = run("ffn wgrad splitK", 256, 2048, 8000, a_mn=True, b_mn=True, bias=False, out=torch.float32, acc=True, splitk=0)
    out["conv2"] = run("conv2 fwd", 160000, 256, 2304, bias=True)
    out["big"] = run("8192^3", 8192, 8192, 8192, bias=False, iters=5)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/r2_gemm_epi_probe.json", "w"), indent=1)
