"""b200st_optimizer_step vs plain torch: Keras Adam, tf.clip_by_value / tf.clip_by_norm per gradient tensor
(neurst/training/gradaccum_keras_model.py:228-233) and the dynamic loss scale state machine
(neurst/training/revised_dynamic_loss_scale.py:60-107)."""
import pytest
import torch

from oracle import restatement as R
from tests import parity_utils as U

pytestmark = pytest.mark.gpu


def _toy_rt(precision="fp32"):
    cfg = dict(R.CONFIGS["speech_transformer_toy"])
    if precision != "fp32":      # the tensor-core precisions need 8-element aligned dims
        cfg = dict(model="speech", d=64, heads=4, enc_layers=1, dec_layers=1, ffn=64, channels=64, feat=80, in_channels=1, vocab=96)
    rt = U.speech_runtime(cfg, precision)
    rt.load_parameters(R.init_params(cfg, seed=1, random_bias=True))
    return rt


@pytest.mark.parametrize("mode", ["value", "norm"])
def test_clipping_matches_tf_semantics(mode):
    rt = _toy_rt()
    g = torch.Generator().manual_seed(0)
    grad = torch.randn(rt.numel, generator=g) * 3.0
    # zero the alignment padding between tensors (the library never writes gradients there)
    keep = torch.zeros(rt.numel, dtype=torch.bool)
    for off, shp in rt.table.values():
        n = 1
        for s_ in shp:
            n *= s_
        keep[off:off + n] = True
    grad = grad * keep
    p0 = rt.params.clone().cpu()
    rt.ensure_grads().copy_(grad.cuda())
    clip = 0.7 if mode == "value" else 2.5
    lr = 1e-2
    rt.adam_step(lr, 1, grad_scale=0.5, zero_grad=True, clip_value=clip if mode == "value" else None,
                 clip_norm=clip if mode == "norm" else None)
    ge = grad * 0.5
    if mode == "value":
        ge = ge.clamp(-clip, clip)
    else:
        for off, shp in rt.table.values():
            n = 1
            for s_ in shp:
                n *= s_
            t = ge[off:off + n]
            ge[off:off + n] = t * clip / max(float(t.norm()), clip)       # tf.clip_by_norm
    p, _, _ = R.adam_update(p0, ge, torch.zeros_like(p0), torch.zeros_like(p0), lr, 1)
    assert float(((rt.params.cpu() - p) * keep).abs().max()) < 1e-6
    assert float(rt.grads.abs().max()) == 0.0


def test_dynamic_loss_scale_skips_and_recovers():
    rt = _toy_rt("fp16")
    st = rt.loss_scale_state
    assert float(st[0]) == 2.0 ** 15
    p0 = rt.params.clone()
    g = torch.randn(rt.numel, device="cuda") * 2.0 ** 15            # "scaled" gradients
    # 1) non-finite gradient: step skipped, scale halved, gradients zeroed, parameters and shadow untouched
    rt.ensure_grads().copy_(g)
    rt.grads[5] = float("inf")
    rt.adam_step(1e-2, 1, growth_steps=2)
    torch.cuda.synchronize()
    assert torch.equal(rt.params, p0) and float(rt.grads.abs().max()) == 0.0
    assert float(st[0]) == 2.0 ** 14 and float(st[2]) == 1.0 and float(st[3]) == 1.0 and float(st[4]) == 0.0
    # 2) finite gradients under the (new) scale: applied with Adam t = 1 and unscaled by 1/S
    S = float(st[0])
    rt.grads.copy_(g)
    rt.adam_step(1e-2, 999, growth_steps=2)       # host step_t is ignored when the device state counts
    torch.cuda.synchronize()
    pe, _, _ = R.adam_update(p0.cpu(), (g / S).cpu(), torch.zeros(rt.numel), torch.zeros(rt.numel), 1e-2, 1)
    assert float((rt.params.cpu() - pe).abs().max()) < 1e-6
    assert float(st[2]) == 0.0 and float(st[4]) == 1.0 and float(st[1]) == 1.0
    assert abs(float(st[5]) - float((g / S).norm())) < 1e-3 * float((g / S).norm())
    # the fp16 shadow follows the master weights
    assert float((rt.shadow.float() - rt.params).abs().max()) < 1e-3 * float(rt.params.abs().max())
    # 3) second finite step reaches growth_steps = 2: scale doubles
    rt.grads.copy_(g)
    rt.adam_step(1e-2, 999, growth_steps=2)
    torch.cuda.synchronize()
    assert float(st[0]) == 2.0 ** 15 and float(st[1]) == 0.0 and float(st[4]) == 2.0
