"""Pins the oracle restatement (oracle/restatement.py) against the reference's own golden vectors:
TF-generated KATs from tests/neurst/** and outputs of the unmodified reference neurst_pt (CPU only)."""
import math

import numpy as np
import torch

from oracle import restatement as R
from tests import golden_utils as G


def sq(a, b):
    return float(((a.double() - G.t(b, torch.float64)) ** 2).sum())


def test_kat_mha_cross():
    # tests/neurst/layers/attentions/multi_head_attention_test.py:7-60
    k = G.load("kat_mha_cross")
    P = {"a.q.kernel": G.t(k["shape:(1, 4)"]), "a.q.bias": G.t(k["shape:(4,)"]),
         "a.kv.kernel": G.t(k["shape:(1, 8)"]), "a.kv.bias": G.t(k["shape:(8,)"]),
         "a.out.kernel": G.t(k["shape:(4, 3)"]), "a.out.bias": G.t(k["shape:(3,)"])}
    out = R.cross_attention(P, "a", G.t(k["var:query"]), G.t(k["var:memory"]), None, 2, R.NO_DROPOUT, "a")
    assert sq(out, k["expect:0"]) < 1e-9


def test_kat_mha_self_bias():
    # multi_head_attention_test.py:63-111
    k = G.load("kat_mha_self")
    P = {"a.qkv.kernel": G.t(k["shape:(2, 12)"]), "a.qkv.bias": G.t(k["shape:(12,)"]),
         "a.out.kernel": G.t(k["shape:(4, 3)"]), "a.out.bias": G.t(k["shape:(3,)"])}
    out = R.self_attention(P, "a", G.t(k["var:query"]), G.t(k["var:bias"]), 2, R.NO_DROPOUT, "a")
    assert sq(out, k["expect:0"]) < 1e-9


def test_kat_mha_self_cache():
    # multi_head_attention_test.py:114-186 (layer called twice with the growing concat cache)
    k = G.load("kat_mha_self_cache")
    P = {"a.qkv.kernel": G.t(k["shape:(2, 12)"]), "a.qkv.bias": G.t(k["shape:(12,)"]),
         "a.out.kernel": G.t(k["shape:(4, 3)"]), "a.out.bias": G.t(k["shape:(3,)"])}
    cache = {"keys": G.t(k["dict:keys"]).reshape(1, 2, 2, 2), "values": G.t(k["dict:values"]).reshape(1, 2, 2, 2)}
    q = G.t(k["var:query"])
    R.self_attention(P, "a", q, None, 2, R.NO_DROPOUT, "a", cache)
    out = R.self_attention(P, "a", q, None, 2, R.NO_DROPOUT, "a", cache)
    assert sq(out, k["expect:0"]) < 1e-9
    assert sq(cache["keys"].reshape(1, 4, 4), k["expect:1"]) < 1e-9
    assert sq(cache["values"].reshape(1, 4, 4), k["expect:2"]) < 1e-9


ENC_CFG = dict(model="none", d=4, heads=2, enc_layers=1, dec_layers=0, ffn=16, vocab=1)
DEC_CFG = dict(model="none", d=4, heads=2, enc_layers=0, dec_layers=1, ffn=16, vocab=1)


def test_kat_encoder():
    # tests/neurst/layers/encoders/transformer_encoder_test.py:21-122 (L2 < 1e-6)
    k = G.load("kat_encoder")
    P = G.kat_params(k, ENC_CFG, "enc")
    out = R.encoder(P, G.t(k["var:inputs"]), G.t(k["var:input_padding"]), 1, 2)
    assert math.sqrt(sq(out, k["expect:0"])) < 1e-6


def test_kat_decoder_train_and_cached_step():
    # tests/neurst/layers/decoders/transformer_decoder_test.py:20-181
    k = G.load("kat_decoder")
    P = G.kat_params(k, DEC_CFG, "dec")
    mem, pad = G.t(k["var:encoder_outputs"]), G.t(k["var:encoder_inputs_padding"])
    out = R.decoder(P, G.t(k["var:decoder_inputs"]), mem, R.input_padding_to_bias(pad), 1, 2)
    assert sq(out, k["expect:0"]) < 1e-9
    caches = [{"keys": torch.zeros(2, 0, 2, 2), "values": torch.zeros(2, 0, 2, 2)}]
    step = R.decoder(P, G.t(k["expect:1"] * 0 + 0).unsqueeze(1) * 0 + G.t(_decoder_step_input()).unsqueeze(1), mem,
                     R.input_padding_to_bias(pad), 1, 2, caches=caches)
    assert sq(step.squeeze(1), k["expect:1"]) < 1e-9
    assert sq(caches[0]["keys"].reshape(2, 1, 4), k["expect:2"]) < 1e-9
    assert sq(caches[0]["values"].reshape(2, 1, 4), k["expect:3"]) < 1e-9


def _decoder_step_input():
    # transformer_decoder_test.py:163-165 (second `decoder_inputs` literal; the extractor keeps the first)
    return [[1.9606155e+00, -1.8318410e+00, -1.8158482e+00, -3.7030798e-01],
            [-1.1357157e-03, 5.5629879e-01, 6.6107117e-02, -1.7330967e+00]]


def test_kat_position():
    # tests/neurst/layers/common_layers_test.py:96-160
    k = G.load("kat_position")
    table = G.t(k["call:set_weights"])
    ids2 = torch.tensor(k["var:inputs2d"], dtype=torch.long)
    ids1 = torch.tensor(k["var:inputs1d"], dtype=torch.long)
    assert sq(R.add_position(table[ids2]), k["expect:0"]) < 1e-9
    assert sq(R.add_position(table[ids1], time=3), k["expect:1"]) < 1e-9


def transformer_toy_params(k, dtype=torch.float32):
    cfg = dict(model="text", d=8, heads=2, enc_layers=2, dec_layers=2, ffn=10, vocab=5, src_vocab=8)
    P = G.kat_params(k, cfg, None, dtype)
    P["trg.emb"] = G.t(k["w:target_symbol_modality/shared/weights"], dtype)
    P["trg.bias"] = G.t(k["w:target_symbol_modality/shared/bias"], dtype)
    P["srcemb.emb"] = G.t(k["w:input_symbol_modality/emb/weights"], dtype)
    return cfg, P


def test_kat_transformer_toy():
    # tests/neurst/models/transformer_test.py:23-665 — full 2+2-layer text Transformer logits (cfg-1 plumbing)
    k = G.load("kat_transformer")
    cfg, P = transformer_toy_params(k)
    logits = R.text_transformer_forward(P, cfg, torch.tensor(k["dict:src"], dtype=torch.long),
                                        G.t(k["dict:src_padding"]), torch.tensor(k["dict:trg_input"], dtype=torch.long))
    assert sq(logits, k["expect:0"]) < 1e-9


def test_refpt_speech_transformer():
    # unmodified reference neurst_pt SpeechTransformer (toy + small): logits, conv front-end, embedded input
    for name in ("refpt_speech_toy", "refpt_speech_small"):
        z, P, cfg = G.refpt_case(name)
        out = R.speech_transformer_forward(P, cfg, G.t(z["src"]), torch.tensor(z["src_length"]),
                                           torch.tensor(z["trg_input"]), return_all=True)
        conv = R.conv_subsample(P, G.t(z["src"]))
        # the reference's own TF<->PT tolerances: L2 < 5e-5 (conv with LN), 5e-6 (model, toy)
        assert float((conv - G.t(z["conv"])).norm()) < 5e-5
        assert float((out["emb"] - G.t(z["emb"])).norm()) < 5e-4
        assert float((out["logits"] - G.t(z["logits"])).abs().max()) < 2e-5, name


def test_label_smoothed_ce_properties():
    # label_smoothed_cross_entropy.py:94-157: unpinned by reference vectors; check the definition directly
    torch.manual_seed(0)
    B, L, V, eps = 3, 5, 11, 0.1
    logits = torch.randn(B, L, V, dtype=torch.float64)
    trg = torch.randint(0, V, (B, L))
    lens = torch.tensor([5, 3, 1])
    nll, ns, nt = R.label_smoothed_ce(logits, trg, lens, eps)
    lp = torch.log_softmax(logits, -1)
    manual = torch.zeros(B, dtype=torch.float64)
    const = -((1 - eps) * math.log(1 - eps) + (V - 1) * (eps / (V - 1)) * math.log(eps / (V - 1) + 1e-20))
    for b in range(B):
        for l in range(int(lens[b])):
            x = 0.0
            for v in range(V):
                tgt = (1 - eps) if v == int(trg[b, l]) else eps / (V - 1)
                x -= tgt * float(lp[b, l, v])
            manual[b] += x - const
    assert torch.allclose(nll, manual, atol=1e-10)
    assert nt.tolist() == [5.0, 3.0, 1.0] and ns.tolist() == [3.0]
    # a perfectly confident, label-smoothing-optimal distribution gives ~0 loss
    opt = torch.log(torch.full((1, 1, V), eps / (V - 1), dtype=torch.float64))
    opt[0, 0, 2] = math.log(1 - eps)
    n2, _, _ = R.label_smoothed_ce(opt, torch.tensor([[2]]), torch.tensor([1]), eps)
    assert abs(float(n2)) < 1e-9


def test_backward_finite_difference_fp64():
    # gradients are unpinned in the reference: check autograd of the restatement against central differences
    cfg = dict(R.CONFIGS["speech_transformer_toy"])
    P = R.init_params(cfg, seed=3, dtype=torch.float64, random_bias=True)
    g = torch.Generator().manual_seed(5)
    src = torch.randn(2, 21, 80, 1, generator=g, dtype=torch.float64)
    lens = torch.tensor([21, 13]); src[1, 13:] = 0
    trg_in = torch.randint(0, cfg["vocab"], (2, 4), generator=g)
    trg = torch.randint(0, cfg["vocab"], (2, 4), generator=g)
    tl = torch.tensor([4, 2])
    for v in P.values():
        v.requires_grad_(True)

    def loss_fn():
        return R.reduce_loss(R.speech_transformer_forward(P, cfg, src, lens, trg_in), trg, tl, 0.1)

    loss = loss_fn()
    grads = torch.autograd.grad(loss, list(P.values()))
    names = list(P.keys())
    rng = np.random.RandomState(0)
    for name in ["src.conv1.kernel", "src.conv2.kernel", "src.ln1.gamma", "src.dense.kernel", "enc.0.att.qkv.kernel",
                 "enc.1.ffn.w2", "dec.0.cross.kv.kernel", "dec.1.self.out.bias", "trg.emb", "trg.bias", "enc.out_ln.beta"]:
        p = P[name]
        ga = grads[names.index(name)]
        for _ in range(2):
            idx = tuple(rng.randint(0, s) for s in p.shape)
            with torch.no_grad():
                old = float(p[idx]); h = 1e-5
                p[idx] = old + h; lp = float(loss_fn())
                p[idx] = old - h; lm = float(loss_fn())
                p[idx] = old
            fd = (lp - lm) / (2 * h)
            assert abs(fd - float(ga[idx])) < 1e-6 * max(1.0, abs(fd)), (name, idx, fd, float(ga[idx]))


def test_noam_and_adam():
    # noam_schedule.py:75-97 evaluated by hand at three steps; Keras Adam epsilon-hat form
    kw = dict(dmodel=256, warmup_steps=25000, initial_factor=3.5, end_factor=1.5, start_decay_at=50000, decay_steps=50000)
    assert abs(R.noam_lr(0, **kw) - 3.5 * 256 ** -0.5 * (1 / 25000) / math.sqrt(25000)) < 1e-15
    assert abs(R.noam_lr(24999, **kw) - 3.5 * 256 ** -0.5 / math.sqrt(25000)) < 1e-12
    assert abs(R.noam_lr(74999, **kw) - 2.5 * 256 ** -0.5 / math.sqrt(75000)) < 1e-12
    p, g = torch.tensor([1.0]), torch.tensor([0.5])
    p1, m1, v1 = R.adam_update(p, g, torch.zeros(1), torch.zeros(1), lr=0.1, t=1)
    m, v = 0.05, 0.02 * 0.25
    assert abs(float(p1) - (1.0 - 0.1 * math.sqrt(1 - 0.98) / (1 - 0.9) * m / (math.sqrt(v) + 1e-9))) < 1e-6
