"""GPU probe for the tcgen05 GEMM (run under gpurun). Each case runs in a subprocess so a trap in one
descriptor hypothesis cannot poison the others. Results -> gpurun_out/gemm_probe.json"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_case(name):
    import torch
    from neurst_b200 import lib
    torch.manual_seed(0)
    dev = "cuda"
    L = lib.load()

    def rnd(*s):
        return torch.randn(*s, device=dev).to(torch.bfloat16)

    def ref_mm(A, B, a_mn, b_mn):
        Af = A.float().transpose(-1, -2) if a_mn else A.float()
        Bf = B.float().transpose(-1, -2) if b_mn else B.float()
        return Af @ Bf.transpose(-1, -2)

    def relerr(c, r):
        return float((c.float() - r).abs().max() / (r.abs().max() + 1e-9))

    out = {}
    kind, *rest = name.split(":")
    if kind == "basic":
        M, N, K, bn = map(int, rest)
        L.b200st_debug_tc(0, 0, 0, 0, bn, 0, 0)
        A, B = rnd(M, K), rnd(N, K)
        Cc = torch.zeros(M, N, device=dev)
        lib.gemm(A, B, Cc)
        torch.cuda.synchronize()
        out["relerr"] = relerr(Cc, ref_mm(A, B, False, False))
    elif kind == "major":
        a_mn, b_mn, lbo, sbo, M, N, K = map(int, rest)
        L.b200st_debug_tc(lbo, sbo, 0, 0, 0, 0, 0)
        A = rnd(K, M) if a_mn else rnd(M, K)
        B = rnd(K, N) if b_mn else rnd(N, K)
        Cc = torch.zeros(M, N, device=dev)
        lib.gemm(A, B, Cc, a_mn=bool(a_mn), b_mn=bool(b_mn))
        torch.cuda.synchronize()
        out["relerr"] = relerr(Cc, ref_mm(A, B, bool(a_mn), bool(b_mn)))
    elif kind == "edge":
        # ragged M/N/K with padded leading dims, bf16 out
        M, N, K = 250, 250, 250
        A = rnd(M, 256)[:, :K]
        B = rnd(N, 256)[:, :K]
        Cf = torch.zeros(M, 256, device=dev)[:, :N]
        lib.gemm(A, B, Cf)
        Cb = torch.zeros(M, 256, device=dev, dtype=torch.bfloat16)[:, :N]
        lib.gemm(A, B, Cb)
        torch.cuda.synchronize()
        r = ref_mm(A, B, False, False)
        out["relerr_f32"] = relerr(Cf, r)
        out["relerr_bf16"] = relerr(Cb, r)
        # MN-major ragged: P[Tq,Tk] @ V[Tk,dh]
        P = rnd(250, 256)[:, :250]
        V = rnd(250, 64)
        O = torch.zeros(250, 64, device=dev)
        lib.gemm(P, V, O, b_mn=True)
        torch.cuda.synchronize()
        out["relerr_pv"] = relerr(O, P.float() @ V.float())
    elif kind == "batched":
        Bsz, T, H, dh = 3, 250, 4, 64
        d = H * dh
        qkv = rnd(Bsz, T, 3 * d)
        q = qkv[:, :, :d].view(Bsz, T, H, dh).permute(0, 2, 1, 3)        # [B,H,T,dh] strided view
        k = qkv[:, :, d:2 * d].view(Bsz, T, H, dh).permute(0, 2, 1, 3)
        v = qkv[:, :, 2 * d:].view(Bsz, T, H, dh).permute(0, 2, 1, 3)
        S = torch.zeros(Bsz, H, T, 256, device=dev)[..., :T]
        lib.gemm(q, k, S, alpha=0.125)
        torch.cuda.synchronize()
        out["relerr_qk"] = relerr(S, 0.125 * (q.float() @ k.float().transpose(-1, -2)))
        P = torch.softmax(S, -1).to(torch.bfloat16)
        Pp = torch.zeros(Bsz, H, T, 256, device=dev, dtype=torch.bfloat16)
        Pp[..., :T] = P
        Pv = Pp[..., :T]
        O = torch.zeros(Bsz, T, H, dh, device=dev, dtype=torch.bfloat16).permute(0, 2, 1, 3)
        lib.gemm(Pv, v, O, b_mn=True)
        torch.cuda.synchronize()
        out["relerr_pv"] = relerr(O, Pv.float() @ v.float())
        # dV = P^T dO  (A MN-major, B MN-major)
        dO = rnd(Bsz, T, H, dh).permute(0, 2, 1, 3)
        dV = torch.zeros(Bsz, H, T, dh, device=dev)
        lib.gemm(Pv, dO, dV, a_mn=True, b_mn=True)
        torch.cuda.synchronize()
        out["relerr_dv"] = relerr(dV, Pv.float().transpose(-1, -2) @ dO.float())
    elif kind == "splitk":
        M, N, K = 256, 2048, 8000
        A, B = rnd(K, M), rnd(K, N)          # wgrad form: both MN-major
        Cc = torch.ones(M, N, device=dev)
        lib.gemm(A, B, Cc, a_mn=True, b_mn=True, accumulate=True, splitk=0)
        torch.cuda.synchronize()
        out["relerr"] = relerr(Cc, 1.0 + A.float().t() @ B.float())
    elif kind == "epi":
        M, N, K = 300, 512, 256
        A, B = rnd(M, K), rnd(N, K)
        bias = torch.randn(N, device=dev)
        res = torch.randn(M, N, device=dev)
        msk = rnd(M, N)
        for tag, kw in [("bias_relu", dict(bias=bias, relu=True)),
                        ("res_drop", dict(bias=bias, residual=res, dropout=(0.1, 1234, 7))),
                        ("mask", dict(mask_src=msk, alpha=0.5))]:
            c1 = torch.zeros(M, N, device=dev)
            c2 = torch.zeros(M, N, device=dev)
            lib.gemm(A, B, c1, **kw)
            lib.gemm(A, B, c2, force_simt=True, **kw)
            torch.cuda.synchronize()
            out[tag] = relerr(c1, c2)
            if tag == "res_drop":
                out["drop_zero_frac"] = float(((c1 - res) == 0).float().mean())
    elif kind == "perf":
        M, N, K = map(int, rest)
        A, B = rnd(M, K), rnd(N, K)
        Cc = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            lib.gemm(A, B, Cc)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        e0.record()
        for _ in range(iters):
            lib.gemm(A, B, Cc)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        out["ms"] = ms
        out["tflops"] = 2.0 * M * N * K / ms / 1e9
        out["relerr"] = relerr(Cc, ref_mm(A, B, False, False))
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(iters):
            torch.matmul(A, B.t(), out=Cc)
        t1.record()
        torch.cuda.synchronize()
        out["cublas_tflops"] = 2.0 * M * N * K / (t0.elapsed_time(t1) / iters) / 1e9
    print("RESULT " + json.dumps(out))


CASES = [
    "basic:128:64:64:64", "basic:256:256:256:0", "basic:256:256:256:64", "basic:256:256:256:128",
    "basic:256:256:256:256", "basic:1024:512:512:0",
    "major:1:0:0:0:256:256:256", "major:0:1:0:0:256:256:256", "major:1:1:0:0:256:256:256",
    "major:1:0:1024:8192:256:256:256", "major:0:1:1024:8192:256:256:256",
    "edge", "batched", "splitk", "epi",
    "perf:8192:2048:256", "perf:8192:256:2048", "perf:8192:8192:8192", "perf:160000:256:2304",
]

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--case":
        run_case(sys.argv[2])
        sys.exit(0)
    cases = sys.argv[1:] or CASES
    results = {}
    for c in cases:
        try:
            r = subprocess.run([sys.executable, __file__, "--case", c], capture_output=True, text=True, timeout=120)
            res = None
            for line in r.stdout.splitlines():
                if line.startswith("RESULT "):
                    res = json.loads(line[7:])
            results[c] = res if res is not None else {"error": (r.stderr or r.stdout)[-600:], "rc": r.returncode}
        except subprocess.TimeoutExpired:
            results[c] = {"error": "timeout"}
        print(c, results[c], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(results, open(os.path.join(ROOT, "gpurun_out", "gemm_probe.json"), "w"), indent=1)
