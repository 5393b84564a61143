"""Host-side data-parallel logic on CPU with the gloo backend, world_size = 2 (the N>1 path of trainer.py):
rank-0 parameter broadcast, per-rank loss normalisation, gradient SUM all-reduce with the 1/world mean folded into
the optimizer, update_cycle accumulation, noam step counting.  The model is a stand-in driven by the oracle."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neurst_b200.trainer import DataParallelTrainer, noam_learning_rate
from oracle import restatement as R


class _Cfg:
    d = 8


class _OracleRuntime:
    """Runtime stand-in with the same surface DataParallelTrainer uses, computing with the CPU oracle."""

    def __init__(self, cfg, seed):
        self.cfg = cfg
        self.config = _Cfg()
        self.P = R.init_params(cfg, seed=seed, dtype=torch.float64, random_bias=True)
        self.names = list(self.P.keys())
        self.params = torch.cat([self.P[k].reshape(-1) for k in self.names])
        self.grads = torch.zeros_like(self.params)
        self.m = torch.zeros_like(self.params)
        self.v = torch.zeros_like(self.params)
        self._shadow_stale = False

    def ensure_grads(self):
        return self.grads

    def unflatten(self):
        out, o = {}, 0
        for k in self.names:
            n = self.P[k].numel()
            out[k] = self.params[o:o + n].view(self.P[k].shape).clone().requires_grad_(True)
            o += n
        return out

    def adam_step(self, lr, step_t, beta1, beta2, eps, grad_scale=1.0, zero_grad=True):
        p, self.m, self.v = R.adam_update(self.params, self.grads * grad_scale, self.m, self.v, lr, step_t, beta1, beta2, eps)
        self.params.copy_(p)
        if zero_grad:
            self.grads.zero_()


class _OracleModel:
    def __init__(self, cfg, seed):
        self.runtime = _OracleRuntime(cfg, seed)
        self.cfg = cfg

    def forward_backward(self, batch, is_training=True, loss_scale=1.0):
        P = self.runtime.unflatten()
        logits = R.speech_transformer_forward(P, self.cfg, batch["src"].double(), batch["src_length"], batch["trg_input"])
        loss = R.reduce_loss(logits, batch["trg"], batch["trg_length"], 0.1) * loss_scale
        g = torch.autograd.grad(loss, list(P.values()))
        self.runtime.grads += torch.cat([x.reshape(-1) for x in g])
        return {"loss": loss.detach()}


def _batch(cfg, seed):
    from tests.parity_utils import synthetic_speech_batch
    return synthetic_speech_batch(cfg, 2, 21, 4, seed=seed)


def _worker(rank, world, port, q):
    torch.set_num_threads(2)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = dict(R.CONFIGS["speech_transformer_toy"])
    model = _OracleModel(cfg, seed=10 + rank)          # different init per rank: broadcast must fix it
    tr = DataParallelTrainer(model, {"beta_1": 0.9, "beta_2": 0.98, "epsilon": 1e-9},
                             dict(dmodel=8, warmup_steps=10, initial_factor=2.0), update_cycle=2)
    tr.broadcast_parameters()
    for micro in range(4):                              # 2 optimizer steps of 2 micro-batches each
        tr.train_step(_batch(cfg, 100 * rank + micro))
    q.put((rank, model.runtime.params.detach().cpu().numpy().copy(), tr.global_step))   # by value: no fd passing after exit
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_data_parallel_matches_manual_average():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in range(2)], key=lambda t: t[0])
    res = [(r, torch.from_numpy(a), n) for r, a, n in res]
    for p in procs:
        p.join(60)
    assert res[0][2] == 2 and res[1][2] == 2
    assert torch.equal(res[0][1], res[1][1])            # replicas stay identical
    # manual reference: rank-0 init, mean over (2 ranks x 2 micro-batches) of per-batch-normalised gradients, Adam + noam
    cfg = dict(R.CONFIGS["speech_transformer_toy"])
    ref = _OracleModel(cfg, seed=10)
    rt = ref.runtime
    for step in range(2):
        for rank in range(2):
            for micro in (2 * step, 2 * step + 1):
                ref.forward_backward(_batch(cfg, 100 * rank + micro))
        lr = noam_learning_rate(step, dmodel=8, warmup_steps=10, initial_factor=2.0)
        rt.adam_step(lr, step + 1, 0.9, 0.98, 1e-9, grad_scale=1.0 / 4.0)
    assert float((rt.params - res[0][1]).abs().max()) < 1e-9


def test_noam_matches_oracle():
    kw = dict(dmodel=256, warmup_steps=25000, initial_factor=3.5, end_factor=1.5, start_decay_at=50000, decay_steps=50000)
    for step in (0, 10, 24999, 25000, 60000, 99999, 200000):
        assert abs(noam_learning_rate(step, **kw) - R.noam_lr(step, **kw)) < 1e-12


def test_throughput_meter_lines_follow_the_reference_format():
    """MetricReductionCallback's `<metric>_per_step` / `<metric>_per_sec` lines (neurst/training/callbacks.py:209-245)."""
    from neurst_b200.trainer import ThroughputMeter
    lines = []
    m = ThroughputMeter(summary_steps=2, world=4, logger=lines.append)
    batch = dict(src=torch.zeros(8, 100, 80, 1), trg_input=torch.zeros(8, 12, dtype=torch.long),
                 src_length=torch.full((8,), 90), trg_length=torch.full((8,), 10))
    m.add(batch); assert m.step_end(1, torch.tensor(3.0), 1e-3) is None
    m.add(batch)
    out = m.step_end(2, torch.tensor(2.5), 1e-3)
    assert out["src_tokens_per_step"] == 8 * 100 * 4 and out["src_real_tokens_per_step"] == 8 * 90 * 4
    assert out["trg_tokens_per_step"] == 8 * 12 * 4 and out["samples_per_step"] == 32
    assert out["src_tokens_per_sec"] > 0 and lines[0].startswith("Update 2\tTrainingLoss=2.50\tSpeed")
