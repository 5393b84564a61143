"""Layer-API incremental decoder step with the reference's dictionary cache (concat mode), on libb200st launches.

Used by layers.TransformerDecoder when called with `cache["decoding_states"]` (the layer-level KATs of the reference's
own tests); the model-level inference path (preallocated caches, device-side greedy search) is neurst_b200/decode.py.


Reference: TransformerDecoder.call with `cache["decoding_states"]` (neurst/layers/decoders/transformer_decoder.py:171-228),
MultiHeadSelfAttention concat cache (multi_head_attention.py:271-276), pre-projected cross-attention memory
(transformer_layers.py:156-160), greedy = beam_size 1 (neurst/layers/search/beam_search.py:254-439).
Every contraction / LayerNorm / softmax is a library kernel launched through the C ABI; this module only sequences
them and keeps the reference's growing [B, i, H, dh] cache tensors.
"""
import torch

from neurst_b200 import lib as L


def _linear(x2d, W, b, relu=False, residual=None):
    out = torch.empty(x2d.shape[0], W.shape[1], dtype=torch.float32, device=x2d.device)
    L.gemm(x2d, W, out, b_mn=True, bias=b, relu=relu, residual=residual)
    return out


def _attend(q, keys, values, H, bias=None):
    """q [B,Tq,u]; keys/values [B,Tk,u] -> ctx [B*Tq,u]"""
    B, Tq, u = q.shape
    Tk, dh = keys.shape[1], u // H
    qh = q.view(B, Tq, H, dh).permute(0, 2, 1, 3)
    kh = keys.view(B, Tk, H, dh).permute(0, 2, 1, 3)
    vh = values.view(B, Tk, H, dh).permute(0, 2, 1, 3)
    S = L.padded_scores(B, H, Tq, Tk, torch.float32, q.device)
    L.gemm(qh, kh, S, alpha=dh ** -0.5)
    P = L.padded_scores(B, H, Tq, Tk, torch.float32, q.device)
    L.softmax(S, P, bias=bias)
    ctx = torch.empty(B, Tq, H, dh, dtype=torch.float32, device=q.device)
    L.gemm(P, vh, ctx.permute(0, 2, 1, 3), b_mn=True)
    return ctx.view(B * Tq, u)


def decoder_step(rt, x, cache, prefix="dec"):
    """x fp32 [B,1,d] (embedded current token); updates cache["decoding_states"] in place; returns [B,1,d]."""
    cfg = rt.config
    P = rt.named_parameters()
    d, H, eps = cfg.d, cfg.heads, cfg.ln_eps
    B, Tq, _ = x.shape
    x2 = x.reshape(B * Tq, d).contiguous()
    mem, mem_bias = cache.get("memory"), cache.get("memory_bias")
    for i in range(cfg.dec_layers):
        lc = cache["decoding_states"]["layer_%d" % i]
        s, c, f = "%s.%d.self" % (prefix, i), "%s.%d.cross" % (prefix, i), "%s.%d.ffn" % (prefix, i)
        h = L.layernorm(x2, P[s + ".ln.gamma"], P[s + ".ln.beta"], eps)
        qkv = _linear(h, P[s + ".qkv.kernel"], P[s + ".qkv.bias"])
        q, k, v = qkv[:, :d], qkv[:, d:2 * d], qkv[:, 2 * d:]
        sa = lc["self_attention"]
        keys = torch.cat([sa["keys"].reshape(B, -1, d), k.reshape(B, Tq, d)], dim=1).contiguous()
        values = torch.cat([sa["values"].reshape(B, -1, d), v.reshape(B, Tq, d)], dim=1).contiguous()
        sa["keys"], sa["values"] = keys.view(B, -1, H, d // H), values.view(B, -1, H, d // H)
        ctx = _attend(q.reshape(B, Tq, d).contiguous(), keys, values, H)
        x2 = _linear(ctx, P[s + ".out.kernel"], P[s + ".out.bias"], residual=x2)
        if mem is not None and cfg.with_cross_attention:
            if "memory" not in lc:      # memorize_memory (transformer_layers.py:156-160): project the memory once
                kv = _linear(mem.reshape(-1, d).contiguous(), P[c + ".kv.kernel"], P[c + ".kv.bias"])
                Tm = mem.shape[1]
                lc["memory"] = {"keys": kv[:, :d].reshape(B, Tm, d).contiguous(),
                                "values": kv[:, d:].reshape(B, Tm, d).contiguous()}
            h = L.layernorm(x2, P[c + ".ln.gamma"], P[c + ".ln.beta"], eps)
            qc = _linear(h, P[c + ".q.kernel"], P[c + ".q.bias"])
            ctx = _attend(qc.view(B, Tq, d), lc["memory"]["keys"], lc["memory"]["values"], H,
                          bias=mem_bias.contiguous() if mem_bias is not None else None)
            x2 = _linear(ctx, P[c + ".out.kernel"], P[c + ".out.bias"], residual=x2)
        h = L.layernorm(x2, P[f + ".ln.gamma"], P[f + ".ln.beta"], eps)
        f1 = _linear(h, P[f + ".w1"], P[f + ".b1"], relu=True)
        x2 = _linear(f1, P[f + ".w2"], P[f + ".b2"], residual=x2)
    out = L.layernorm(x2, P[prefix + ".out_ln.gamma"], P[prefix + ".out_ln.beta"], eps)
    return out.view(B, Tq, d)
