"""`neurst-run`-style entries over the C-ABI step loop (SURVEY.md 8b iii): a thin `trainer` / `predict` pair that consumes
the reference's yaml configuration for the speech-to-text path and drives libb200st.

    python -m neurst_b200.cli --config_paths train.yml[,more.yml] [--model_dir DIR] [--hparams_set speech_transformer_s]
                              [--entry.class predict] [--train_steps 1000] [--task.params.max_src_len 2000] ...
    python -m torch.distributed.run --nproc-per-node 8 -m neurst_b200.cli --config_paths train.yml      # data parallel

Host mirror of
  neurst/cli/run_exp.py + neurst/utils/flags_core.py:207-330   config files merged in order, command line on top
  neurst/exps/trainer.py:180-315 (Trainer.run)                 model + optimizer + schedule + dataset -> fit loop, checkpoints
  neurst/exps/sequence_generator.py:62-137                     restore -> search -> hypotheses file
  neurst/tasks/speech2text.py                                  task parameters (data side: neurst_b200/data.py)
  neurst/data/data_pipelines/text_data_pipeline.py:83-93       target vocabulary meta: tokens + <UNK>, <SEQ_BEG>, <SEQ_END>

Keys understood (the ones of examples/speech_transformer/*/st_training_args.yml): entry.class / entry.params {train_steps,
summary_steps, save_checkpoint_steps, update_cycle, clip_value, clip_norm, criterion.params.label_smoothing, optimizer.params,
lr_schedule.params, pretrain_model}, dataset.class AudioTFRecordDataset / dataset.params, task.class SpeechToText /
task.params, model.class / model.params / hparams_set, model_dir, dtype.  Transcripts must already be token ids (the
reference's `create_tfrecords` with a projected transcript): its tokenizers / BPE are outside this path and raw text is
refused with that message.  Checkpoints are NumPy archives keyed by the reference's variable names (checkpoints.py).
"""
import copy
import glob
import json
import logging
import os
import re
import sys

import torch

LOG = logging.getLogger("neurst_b200")

SECTIONS = ("entry.params", "task.params", "dataset.params", "model.params")


# ------------------------------------------------------------------------------------------------ configuration
def deep_merge(base, new):
    """Later files win; `*.params` dictionaries are merged key by key (flags_core.py: deep_merge_dict)."""
    out = copy.deepcopy(base)
    for k, v in (new or {}).items():
        if isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = deep_merge(out[k], v)
        else:
            out[k] = copy.deepcopy(v)
    return out


def _coerce(text):
    import yaml
    try:
        return yaml.safe_load(text)
    except Exception:
        return text


def parse_command_line(argv):
    """--config_paths a,b  plus any number of `--key value` overrides.  A dotted key addresses the nested dictionary
    (`--task.params.max_src_len 2000`, `--entry.class predict`); a plain key is matched against the parameter sections, as
    the reference's flat flags are (`--train_steps 10` -> entry.params.train_steps)."""
    paths, overrides = [], []
    i = 0
    while i < len(argv):
        a = argv[i]
        if not a.startswith("--"):
            raise SystemExit("unexpected argument %r" % a)
        key = a[2:]
        if "=" in key:
            key, val = key.split("=", 1)
        else:
            i += 1
            if i >= len(argv):
                raise SystemExit("flag --%s needs a value" % key)
            val = argv[i]
        if key == "config_paths":
            paths.extend(p for p in val.split(",") if p)
        else:
            overrides.append((key, _coerce(val)))
        i += 1
    return paths, overrides


def _set_dotted(cfg, key, val):
    # "a.params.b.c" -> cfg["a.params"]["b.c"]; "entry.class" -> cfg["entry.class"]
    m = re.match(r"^([a-z_]+\.params)\.(.+)$", key)
    if m:
        cfg.setdefault(m.group(1), {})[m.group(2)] = val
    else:
        cfg[key] = val


def load_config(paths, overrides=()):
    import yaml
    cfg = {}
    for p in paths:
        with open(p) as f:
            cfg = deep_merge(cfg, yaml.safe_load(f) or {})
    for key, val in overrides:
        if "." in key or key in ("model_dir", "hparams_set", "dtype", "output_file"):
            _set_dotted(cfg, key, val)
            continue
        hits = [s for s in SECTIONS if isinstance(cfg.get(s), dict) and key in cfg[s]]
        if len(hits) > 1:
            raise SystemExit("flag --%s is ambiguous (%s): use the dotted form" % (key, ", ".join(hits)))
        cfg.setdefault(hits[0] if hits else "entry.params", {})[key] = val
    return cfg


def target_meta(task_params):
    """Vocabulary meta of the target side (text_data_pipeline.py:83-93, data/text/vocab.py): ids 0..n-1 are the file's
    tokens, then <UNK>, <SEQ_BEG>, <SEQ_END>; padding = EOS."""
    tp = task_params
    n = tp.get("vocab_size")
    tokens = None
    path = None
    for k in ("transcript_data_pipeline.params", "trg_data_pipeline.params", "translation_data_pipeline.params"):
        if isinstance(tp.get(k), dict) and tp[k].get("vocab_path"):
            path = tp[k]["vocab_path"]
    if path and os.path.exists(path):
        with open(path, encoding="utf-8") as f:
            tokens = [l.rstrip("\n").split("\t")[0].split(" ")[0] for l in f if l.strip()]
        n = len(tokens) + 3
    if n is None:
        raise SystemExit("the target vocabulary is unknown: give task.params.vocab_size or a readable *_data_pipeline.params.vocab_path")
    n = int(n)
    return {"vocab_size": n, "unk_id": n - 3, "bos_id": n - 2, "eos_id": n - 1, "pad_id": n - 1, "padding_mode": "eos_as_padding",
            "tokens": tokens}


def resolve(cfg, world=1):
    """Everything the entries need, as plain dictionaries (also what the CPU test checks)."""
    from neurst_b200.models import speech_transformer_hparams
    ep, tp = dict(cfg.get("entry.params") or {}), dict(cfg.get("task.params") or {})
    task_cls = str(cfg.get("task.class", "SpeechToText")).lower().replace("_", "")
    if task_cls not in ("speechtotext", "speech2text"):
        raise SystemExit("task.class %r is outside this path (SpeechToText only)" % cfg.get("task.class"))
    ds_cls = str(cfg.get("dataset.class", "AudioTFRecordDataset")).lower().replace("_", "")
    if ds_cls not in ("audiotfrecorddataset", "audiotfrecord"):
        raise SystemExit("dataset.class %r is outside this path (AudioTFRecordDataset only)" % cfg.get("dataset.class"))
    hp = speech_transformer_hparams(cfg.get("hparams_set") or "speech_transformer_s")
    if hp is None:
        raise SystemExit("unknown hparams_set %r" % cfg.get("hparams_set"))
    model_cls = str(cfg.get("model.class", hp["model.class"])).lower().replace("_", "")
    if model_cls not in ("speechtransformer", "b200speechtransformer", "b200st"):
        raise SystemExit("model.class %r is outside this path (SpeechTransformer only)" % cfg.get("model.class"))
    model_params = dict(hp["model.params"])
    model_params.update(cfg.get("model.params") or {})
    opt = dict(hp["optimizer.params"]); opt.update(ep.get("optimizer.params") or {})
    lr = dict(hp["lr_schedule.params"]); lr.update(ep.get("lr_schedule.params") or {})
    if str(ep.get("lr_schedule.class", "noam")).lower() != "noam" or str(ep.get("optimizer.class", "adam")).lower() != "adam":
        raise SystemExit("this path implements the presets' Adam + noam schedule")
    crit = ep.get("criterion.params") or {}
    # batch_size counts frames when batch_by_frames (speech2text.py:296-310): per-GPU budget = batch_size // replicas
    per_gpu = tp.get("batch_size_per_gpu")
    if per_gpu is None:
        if tp.get("batch_size") is None:
            raise SystemExit("task.params needs batch_size or batch_size_per_gpu (frames)")
        per_gpu = int(tp["batch_size"]) // max(1, world)
    task = dict(max_src_len=tp.get("max_src_len"), max_trg_len=tp.get("max_trg_len"), batch_size_per_gpu=int(per_gpu),
                audio_feature_dim=int(tp.get("audio_feature_dim", 80)), audio_feature_channels=int(tp.get("audio_feature_channels", 1)),
                truncate_src=bool(tp.get("truncate_src", False)), truncate_trg=bool(tp.get("truncate_trg", False)),
                min_src_bucket_boundary=tp.get("min_src_bucket_boundary", 128),
                frame_transcript_ratio=tp.get("experimental_frame_transcript_ratio"),
                disable_batch_efficiency=bool(tp.get("disable_batch_efficiency", False)), specaug=tp.get("specaug"))
    dsp = dict(cfg.get("dataset.params") or {})
    dataset = {"data_path": dsp.get("data_path"), "feature_key": dsp.get("feature_key", "audio"),
               "transcript_key": dsp.get("transcript_key", "transcript")}
    return {
        "entry": str(cfg.get("entry.class", "trainer")).lower(),
        "model_dir": cfg.get("model_dir") or ep.get("model_dir") or "./b200st_model",
        "precision": {"float16": "fp16", "fp16": "fp16", "bfloat16": "bf16", "bf16": "bf16", "float32": "fp32", "fp32": "fp32"}[
            str(cfg.get("dtype", "float16")).lower()],
        "model_params": model_params, "optimizer": opt, "lr_schedule": lr,
        "label_smoothing": float(crit.get("label_smoothing", 0.1)),
        "train_steps": int(ep.get("train_steps", 10000000)), "summary_steps": int(ep.get("summary_steps", 200)),
        "save_checkpoint_steps": int(ep.get("save_checkpoint_steps", 1000)), "update_cycle": int(ep.get("update_cycle", 1)),
        "clip_value": ep.get("clip_value"), "clip_norm": ep.get("clip_norm"), "pretrain_model": ep.get("pretrain_model"),
        "seed": int(ep.get("random_seed", 1234) or 1234), "max_to_keep": int(ep.get("checkpoints_max_to_keep", 8)),
        "task": task, "dataset": dataset, "trg_meta": target_meta(tp),
        "search": {"maximum_decode_length": int(ep.get("maximum_decode_length", tp.get("max_trg_len") or 256)),
                   "extra_decode_length": int(ep.get("extra_decode_length", 50)),
                   "minimum_decode_length": int(ep.get("minimum_decode_length", 0))},
        "output_file": cfg.get("output_file") or ep.get("output_file"),
    }


# ------------------------------------------------------------------------------------------------ checkpoints
def latest_checkpoint(model_dir):
    """(path, step) of the newest `ckpt-<step>.npz` under model_dir (tf.train.latest_checkpoint's role), or (None, 0)."""
    best = (None, 0)
    for p in glob.glob(os.path.join(model_dir, "ckpt-*.npz")):
        m = re.search(r"ckpt-(\d+)\.npz$", p)
        if m and int(m.group(1)) >= best[1]:
            best = (p, int(m.group(1)))
    return best


def save_checkpoint(rt, model_dir, step, max_to_keep=8):
    from neurst_b200 import checkpoints as CK
    os.makedirs(model_dir, exist_ok=True)
    path = os.path.join(model_dir, "ckpt-%d.npz" % step)
    tmp = path + ".tmp.npz"
    CK.save_npz(rt, tmp)
    os.replace(tmp, path)
    kept = sorted((int(re.search(r"ckpt-(\d+)\.npz$", p).group(1)), p) for p in glob.glob(os.path.join(model_dir, "ckpt-*.npz")))
    for _, p in kept[:-max_to_keep] if max_to_keep > 0 else []:
        os.remove(p)
    with open(os.path.join(model_dir, "checkpoint"), "w") as f:          # the index file the reference's managers keep
        json.dump({"model_checkpoint_path": os.path.basename(path), "all_model_checkpoint_paths": [os.path.basename(p) for _, p in kept[-max_to_keep:]]}, f)
    return path


# ------------------------------------------------------------------------------------------------ entries
def _dist_setup():
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("neurst_b200.cli needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1 and not torch.distributed.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world, rank, local


def _build_model(plan):
    from neurst_b200.models import SpeechTransformer
    tm = {k: v for k, v in plan["trg_meta"].items() if k != "tokens"}
    if plan["precision"] != "fp32" and tm["vocab_size"] % 8 != 0:
        raise SystemExit("the 16-bit tensor-core path needs a target vocabulary size that is a multiple of 8 (TMA row alignment of the "
                         "logits gradient); this vocabulary has %d entries (tokens + <UNK>, <SEQ_BEG>, <SEQ_END>): pad the vocabulary file "
                         "with %d unused tokens or run with --dtype float32" % (tm["vocab_size"], -tm["vocab_size"] % 8))
    src_meta = {"audio_feature_dim": plan["task"]["audio_feature_dim"], "audio_feature_channels": plan["task"]["audio_feature_channels"]}
    return SpeechTransformer.new(plan["model_params"], src_meta, tm, precision=plan["precision"], label_smoothing=plan["label_smoothing"])


def _examples(plan, rank, world, training):
    """Endless (training) or single-pass stream of preprocessed samples of this rank's file shard."""
    from neurst_b200.data import SpeechToText
    from neurst_b200.tfrecord import AudioTFRecordDataset
    ds = AudioTFRecordDataset(plan["dataset"])
    tk = dict(plan["task"])
    task = SpeechToText({k: v for k, v in plan["trg_meta"].items() if k != "tokens"}, world=1,
                        padding_mode=plan["trg_meta"]["padding_mode"], **tk)
    proc = task.preprocess_fn(ds.status, training=training, with_label=training)
    from neurst_b200.tfrecord import glob_tfrecords
    shards = world if len(glob_tfrecords(plan["dataset"]["data_path"])) >= world else 1     # fewer files than ranks: every rank reads all

    def stream():
        while True:
            n = 0
            for ex in ds.build_iterator(map_func=proc, shard_id=rank if shards > 1 else 0, total_shards=shards)():
                n += 1
                yield ex
            if not training or n == 0:
                return
    return task, stream()


def train(plan):
    from neurst_b200 import checkpoints as CK
    from neurst_b200.trainer import DataParallelTrainer, HostPipeline
    world, rank, local = _dist_setup()
    model = _build_model(plan)
    start_path, start_step = latest_checkpoint(plan["model_dir"])
    if start_path:
        CK.load_npz(model.runtime, start_path)
        LOG.info("restored %s", start_path)
    elif plan["pretrain_model"]:
        src = plan["pretrain_model"][0] if isinstance(plan["pretrain_model"], (list, tuple)) else plan["pretrain_model"]
        src = latest_checkpoint(src)[0] if os.path.isdir(src) else src
        CK.load_npz(model.runtime, src, strict=False)
        LOG.info("initialised from %s", src)
    else:
        model.init_parameters(plan["seed"])
    trainer = DataParallelTrainer(model, plan["optimizer"], plan["lr_schedule"], update_cycle=plan["update_cycle"],
                                  use_cuda_graph=True, clip_value=plan["clip_value"], clip_norm=plan["clip_norm"],
                                  summary_steps=plan["summary_steps"] if rank == 0 else 0, logger=LOG.info)
    trainer.global_step = start_step
    trainer.broadcast_parameters()
    task, examples = _examples(plan, rank, world, training=True)
    gen = torch.Generator().manual_seed(plan["seed"] + rank)
    from neurst_b200.data import Prefetcher
    import itertools
    last = None
    micro_target = (plan["train_steps"] - start_step) * plan["update_cycle"]
    if micro_target > 0:
        # reading, parsing, padding and pinning of the next batches run in a background thread (tf.data prefetch's role)
        batches = Prefetcher(itertools.islice((per_rank[0] for per_rank in task.train_batches(examples, generator=gen, pin=True)),
                                              micro_target), depth=4, init=lambda: torch.cuda.set_device(local))
        try:
            losses = HostPipeline(trainer).run(batches, seed0=start_step * plan["update_cycle"] + 1)
            saved_at = start_step
            for loss in losses:
                if loss is not None:
                    last = loss
                step = trainer.global_step
                if rank == 0 and step > saved_at and step % plan["save_checkpoint_steps"] == 0:
                    torch.cuda.synchronize()
                    save_checkpoint(model.runtime, plan["model_dir"], step, plan["max_to_keep"])
                    saved_at = step
        finally:
            batches.close()
    torch.cuda.synchronize()
    if rank == 0 and trainer.global_step > start_step:
        save_checkpoint(model.runtime, plan["model_dir"], trainer.global_step, plan["max_to_keep"])
    if world > 1:
        trainer.close()
        torch.distributed.barrier()
    return {"global_step": trainer.global_step, "loss": last}


def ids_to_text(ids, eos_id, tokens):
    out = []
    for i in ids:
        if i == eos_id:
            break
        out.append(int(i))
    if tokens is None:
        return " ".join(str(i) for i in out)
    words = [tokens[i] if i < len(tokens) else "<UNK>" for i in out]
    return " ".join(words).replace("@@ ", "")                            # BPE continuation marks (subword-nmt convention)


def predict(plan):
    from neurst_b200 import checkpoints as CK
    from neurst_b200.decode import MAX_ROWS
    world, rank, _ = _dist_setup()
    model = _build_model(plan)
    path, step = latest_checkpoint(plan["model_dir"])
    if not path:
        raise SystemExit("no ckpt-*.npz under %s" % plan["model_dir"])
    CK.load_npz(model.runtime, path)
    task, examples = _examples(plan, rank, world, training=False)
    tm = plan["trg_meta"]
    dim = plan["task"]["audio_feature_dim"] * plan["task"]["audio_feature_channels"]
    hyps = []
    group = []

    def flush():
        if not group:
            return
        T = max(int(e["audio_length"]) for e in group)
        src = torch.zeros(len(group), T, dim)
        for j, e in enumerate(group):
            src[j, :int(e["audio_length"])] = e["audio"]
        batch = task.example_to_input({"audio": src, "audio_length": torch.tensor([int(e["audio_length"]) for e in group])}, infer=True)
        ids, logprob = model.greedy_search({"src": batch["src"], "src_length": batch["src_length"]}, **plan["search"])
        for row, lp in zip(ids.cpu().tolist(), logprob.cpu().tolist()):
            hyps.append((ids_to_text(row, tm["eos_id"], tm.get("tokens")), lp))
        group.clear()

    for ex in examples:
        group.append(ex)
        if len(group) == MAX_ROWS:
            flush()
    flush()
    out = plan["output_file"]
    if out and rank == 0:
        with open(out, "w", encoding="utf-8") as f:
            for text, _ in hyps:
                f.write(text + "\n")
    return {"checkpoint_step": step, "hypotheses": hyps}


def main(argv=None):
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(message)s")
    paths, overrides = parse_command_line(list(sys.argv[1:] if argv is None else argv))
    cfg = load_config(paths, overrides)
    plan = resolve(cfg, world=int(os.environ.get("WORLD_SIZE", "1")))
    if plan["entry"] in ("trainer", "train"):
        res = train(plan)
    elif plan["entry"] in ("predict", "sequencegenerator", "sequence_generator"):
        res = predict(plan)
        res = {"checkpoint_step": res["checkpoint_step"], "n_hypotheses": len(res["hypotheses"])}
    else:
        raise SystemExit("entry.class %r is outside this path (trainer / predict)" % plan["entry"])
    LOG.info("%s", json.dumps(res))
    return res


if __name__ == "__main__":
    main()
