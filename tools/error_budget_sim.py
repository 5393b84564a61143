"""CPU simulation of the operand-rounding error budget of the tensor-core path (dev tool, not product).

Runs the oracle forward of speech_transformer_s in fp64 and again with every MMA operand / stored activation rounded
to bf16 or fp16 at the sites where the CUDA path stores 16-bit values, and prints max-abs / rms logits error.
Usage: python tools/error_budget_sim.py [B T L]
"""
import sys, math, torch
sys.path.insert(0, ".")
from oracle import restatement as R

def make_q(kind):
    if kind == "none":
        return lambda x: x
    dt = torch.bfloat16 if kind == "bf16" else torch.float16
    return lambda x: x.to(dt).to(x.dtype)

def fwd(P, cfg, src, src_length, trg_input, q, qlog=None, split_logits=False, stages=None):
    H = cfg["heads"]; eps = 1e-6
    def lin(x, w, b): return q(x) @ q(P[w]) + P[b]
    def mha(pre, x, mem, bias, causal):
        d = x.shape[-1]
        if mem is None:
            qkv = q(lin(x, pre + ".qkv.kernel", pre + ".qkv.bias")); qq, k, v = qkv[..., :d], qkv[..., d:2*d], qkv[..., 2*d:]
        else:
            qq = q(lin(x, pre + ".q.kernel", pre + ".q.bias")); kv = q(lin(mem, pre + ".kv.kernel", pre + ".kv.bias")); k, v = kv[..., :d], kv[..., d:]
        dh = d // H
        B, Tq, Tk = qq.shape[0], qq.shape[1], k.shape[1]
        qh = qq.reshape(B, Tq, H, dh).permute(0, 2, 1, 3); kh = k.reshape(B, Tk, H, dh).permute(0, 2, 1, 3); vh = v.reshape(B, Tk, H, dh).permute(0, 2, 1, 3)
        S = (qh @ kh.transpose(-1, -2)) * dh ** -0.5
        if bias is not None: S = S + bias[:, None, None, :]
        if causal: S = S + R.lower_triangle_attention_bias(Tq, S.dtype)
        Pm = torch.softmax(S, -1)
        # fused kernel: unnormalised exp2 probabilities are the bf16 MMA operand, normalised by the fp32 row sum after PV
        m = S.max(-1, keepdim=True).values; e = torch.exp(S - m); l = e.sum(-1, keepdim=True)
        o = (q(e) @ vh) / l
        o = q(o.permute(0, 2, 1, 3).reshape(B, Tq, d))
        return lin(o, pre + ".out.kernel", pre + ".out.bias")
    def ln(x, pre): return R.layer_norm(x, P[pre + ".gamma"], P[pre + ".beta"], eps)
    def ffn(pre, x):
        h = q(torch.relu(lin(x, pre + ".w1", pre + ".b1")))
        return lin(h, pre + ".w2", pre + ".b2")
    # front-end
    x = src.permute(0, 3, 1, 2)
    w1 = P["src.conv1.kernel"].permute(3, 2, 0, 1)
    z1 = torch.nn.functional.conv2d(x, w1, P["src.conv1.bias"], stride=2, padding=1).permute(0, 2, 3, 1)
    mean = z1.mean(-1, keepdim=True); var = ((z1 - mean) ** 2).mean(-1, keepdim=True)
    xhat = q((z1 - mean) / torch.sqrt(var + eps))
    y1 = q(torch.relu(q(xhat * q(P["src.ln1.gamma"])) + q(P["src.ln1.beta"])))     # packed 16-bit affine in im2col
    w2 = q(P["src.conv2.kernel"]).permute(3, 2, 0, 1)
    z2 = q(torch.nn.functional.conv2d(y1.permute(0, 3, 1, 2), w2, P["src.conv2.bias"], stride=2, padding=1).permute(0, 2, 3, 1))
    y2 = q(torch.relu(R.layer_norm(z2, P["src.ln2.gamma"], P["src.ln2.beta"], eps)))
    B, T2, F2, C = y2.shape
    e0 = y2.reshape(B, T2, F2 * C) @ q(P["src.dense.kernel"]) + P["src.dense.bias"]
    emb = R.add_position(e0)
    if stages is not None: stages["emb"] = emb
    padding = R.input_length_to_padding(R.length_after_conv(src_length), T2, emb.dtype)
    bias = R.input_padding_to_bias(padding)
    x = emb
    for i in range(cfg["enc_layers"]):
        a, f = "enc.%d.att" % i, "enc.%d.ffn" % i
        x = x + mha(a, ln(x, a + ".ln"), None, bias, False)
        x = x + ffn(f, ln(x, f + ".ln"))
        if stages is not None: stages["enc%d" % i] = x
    enc = q(ln(x, "enc.out_ln"))
    if stages is not None: stages["enc"] = enc
    y = R.target_embed(P, trg_input)
    for i in range(cfg["dec_layers"]):
        s, c, f = "dec.%d.self" % i, "dec.%d.cross" % i, "dec.%d.ffn" % i
        y = y + mha(s, ln(y, s + ".ln"), None, None, True)
        y = y + mha(c, ln(y, c + ".ln"), enc, bias, False)
        y = y + ffn(f, ln(y, f + ".ln"))
    dec = ln(y, "dec.out_ln")
    if stages is not None: stages["dec"] = dec
    ql = qlog or q
    E = P["trg.emb"]
    if split_logits:
        dh_, eh = ql(dec), ql(E); dl, el = ql(dec - dh_), ql(E - eh)
        lg = dh_ @ eh.t() + dh_ @ el.t() + dl @ eh.t()
    else:
        lg = ql(dec) @ ql(E).t()
    return lg + P["trg.bias"]

def main():
    B, T, L = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (2, 200, 24)
    cfg = R.CONFIGS["speech_transformer_s"]
    P = R.init_params(cfg, seed=0, dtype=torch.float64, random_bias=True)
    g = torch.Generator().manual_seed(1)
    src = torch.randn(B, T, 80, 1, generator=g, dtype=torch.float64)
    sl = torch.tensor([T] + [max(T * 3 // 4, 1)] * (B - 1)); ti = torch.randint(4, cfg["vocab"], (B, L), generator=g)
    st0 = {}
    ref = fwd(P, cfg, src, sl, ti, make_q("none"), stages=st0)
    ref2 = R.speech_transformer_forward(P, cfg, src, sl, ti)
    print("sim(no rounding) vs oracle: %.2e" % (ref - ref2).abs().max().item())
    for kind in ("bf16", "fp16"):
        for split in (False, True):
            st = {}
            out = fwd(P, cfg, src, sl, ti, make_q(kind), split_logits=split, stages=st)
            e = out - ref
            print("%s split_logits=%d: logits max-abs %.3e rms %.3e | " % (kind, split, e.abs().max().item(), e.pow(2).mean().sqrt().item()) +
                  " ".join("%s %.1e" % (k, (st[k] - st0[k]).abs().max().item() / st0[k].abs().max().item()) for k in ("emb", "enc0", "enc5", "enc11", "enc", "dec")))
    out = fwd(P, cfg, src, sl, ti, make_q("bf16"), qlog=make_q("fp16"))
    print("bf16 body + fp16 logits: max-abs %.3e" % (out - ref).abs().max().item())

if __name__ == "__main__":
    torch.set_num_threads(16)
    main()
