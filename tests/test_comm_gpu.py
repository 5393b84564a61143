"""In-library NCCL gradient aggregation (b200st_comm_* / b200st_train_step): two ranks on two GPUs.  Each rank computes
its local gradients twice with the same dropout seed — once without the all-reduce, once with the bucketed, overlapped
all-reduce inside the backward pass — and the reduced arena must equal the sum of the ranks' local arenas (gathered
through torch.distributed).  Also the eager vs CUDA-graph step and the rank-0 parameter broadcast."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from neurst_b200.trainer import build_speech_transformer_trainer, synthetic_batch
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        tr, _ = build_speech_transformer_trainer("speech_transformer_s", vocab_size=96, precision="fp16", label_smoothing=0.1,
                                                 seed=5 + rank, use_cuda_graph=False)       # different init per rank
        rt = tr.rt
        assert tr.lib_comm and rt.comm_stats()["world"] == world
        p0 = rt.params.clone()
        g0 = [torch.empty_like(p0) for _ in range(world)]
        dist.all_gather(g0, p0)
        assert torch.equal(g0[0], p0), "broadcast_parameters: every rank must hold rank 0's parameters"
        batch = synthetic_batch(4, 160, 12, 96, seed=100 + rank, device="cuda")
        b = dict(batch); b.update(training=True, seed=77 * world + rank, want_logits=False)
        rt.ensure_grads().zero_()
        rt.run(b, backward=True)
        local = rt.grads.clone()
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        expect = sum(gathered)
        rt.grads.zero_()
        b2 = dict(b); b2["allreduce"] = True
        rt.run(b2, backward=True)
        torch.cuda.synchronize()
        st = rt.comm_stats()
        err = float((rt.grads - expect).norm() / expect.norm())
        # graph-captured step with the all-reduce inside the graph
        from neurst_b200.runtime import GraphedTrainStep
        gs = GraphedTrainStep(rt, 4, 160, 12, allreduce=True).capture()
        rt.grads.zero_()
        gs(batch, 77 * world + rank)
        torch.cuda.synchronize()
        err_g = float((rt.grads - expect).norm() / expect.norm())
        # per-bucket error and the run-to-run noise floor of the local gradients (same batch, same seed, no all-reduce)
        rt.grads.zero_()
        rt.run(b, backward=True)
        torch.cuda.synchronize()
        noise = float((rt.grads - local).norm() / local.norm())
        q.put((rank, err, err_g, st["reduced_elems"], st["calls"], rt.numel, noise))
        del gs              # a live graph that captured the all-reduces keeps ncclCommDestroy waiting
        tr.close()
    finally:
        dist.barrier()
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_bucketed_allreduce_inside_the_library_two_ranks():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(2)]
    hung = False
    for p in procs:
        p.join(30)
        if p.is_alive():
            p.terminate()
            hung = True
    for rank, err, err_g, n_red, calls, numel, noise in res:
        print("rank %d: reduced-vs-sum rel err eager %.2e graph %.2e, run-to-run noise of the local gradients %.2e" % (rank, err, err_g, noise))
        # the reduced arena is the sum of two independently recomputed local arenas: the bound is the run-to-run noise of the
        # backward pass itself (fp32 atomics order -> 1-ulp flips of 16-bit intermediates), not of the reduction
        assert err < max(1e-5, 4 * noise) and err_g < max(1e-5, 4 * noise), (rank, err, err_g, noise)
        assert n_red == numel and calls == 4, (n_red, numel, calls)   # whole arena, in 4 buckets
    assert not hung, "a rank did not leave after closing the communicator"
