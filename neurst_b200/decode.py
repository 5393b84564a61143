"""Inference on top of libb200st: encoder pass, decoding cache, cached decoder step, greedy search.

Host mirror of
  EncoderDecoderModel.get_symbols_to_logits_fn      neurst/models/encoder_decoder_model.py:211-261
  TransformerDecoder.create_decoding_internal_cache neurst/layers/decoders/transformer_decoder.py:105-147
  sequence_beam_search (beam_size = 1)              neurst/layers/search/beam_search.py:254-439
  SequenceGenerator                                 neurst/exps/sequence_generator.py:62-86

Everything numeric is a library call (b200st_encode / b200st_decode_init / b200st_decode_step / b200st_greedy_search):
the key/value caches are preallocated device buffers owned by `DecodingCache`, the token ids and the position live on the
device, and the greedy loop replays one captured CUDA graph per token inside the library.
"""
import ctypes as C

import torch

from neurst_b200 import lib as L

MAX_ROWS = 8      # rows (batch x beam) per decoding step supported by the decode kernels


class DecodingCache:
    """`cache` argument of symbols_to_logits_fn: encoder memory (pre-projected per layer) + self-attention K/V."""

    def __init__(self, rt, B, Tm, max_len, memory_bias, use_shadow=False):
        cfg = rt.config
        if not 1 <= B <= MAX_ROWS:
            raise L.B200STError("decoding supports 1..%d rows per step, got %d" % (MAX_ROWS, B))
        dev = rt.device
        self.rt, self.B, self.Tm, self.max_len = rt, B, Tm, max_len
        self.cross_kv = torch.zeros(cfg.dec_layers, B, max(Tm, 1), 2 * cfg.d, dtype=torch.float32, device=dev)
        self.self_kv = torch.zeros(cfg.dec_layers, 2, B, max_len, cfg.d, dtype=torch.float32, device=dev)
        self.memory_bias = memory_bias
        n = int(rt.lib.b200st_decode_scratch_floats(rt.handle, B))
        self.scratch = torch.zeros(n, dtype=torch.float32, device=dev)
        self.time = torch.zeros(1, dtype=torch.int32, device=dev)
        st = L.DecodeState()
        st.B, st.Tm, st.max_len = B, Tm, max_len
        st.cross_kv, st.self_kv = self.cross_kv.data_ptr(), self.self_kv.data_ptr()
        st.memory_bias = memory_bias.data_ptr() if memory_bias is not None else None
        st.scratch = self.scratch.data_ptr()
        st.use_shadow = int(bool(use_shadow))
        self.state = st

    # reference-shaped view of the cache (transformer_decoder.py:119-147) for inspection / tests
    def as_dict(self, upto):
        d, H = self.rt.config.d, self.rt.config.heads
        out = {}
        for i in range(self.rt.config.dec_layers):
            out["layer_%d" % i] = {
                "self_attention": {"keys": self.self_kv[i, 0, :, :upto].reshape(self.B, upto, H, d // H),
                                   "values": self.self_kv[i, 1, :, :upto].reshape(self.B, upto, H, d // H)},
                "memory": {"keys": self.cross_kv[i, :, :, :d], "values": self.cross_kv[i, :, :, d:]}}
        return out


def _bufs(rt, need=0):
    if rt._shadow_stale:
        rt.refresh_shadow()
    return rt._buffers(max(int(need), 256), False)


def encode(rt, inputs):
    """Modality + encoder: returns (encoder_outputs fp32 [B,T',d], memory_bias fp32 [B,T'] additive)."""
    cfg = rt.config
    bt = L.Batch()
    keep = []

    def put(field, t, dtype):
        t = torch.as_tensor(t).to(device=rt.device, dtype=dtype).contiguous()
        keep.append(t)
        setattr(bt, field, t.data_ptr())

    src = inputs["src"]
    B, T = src.shape[0], src.shape[1]
    if cfg.model_type == L.MODEL_SPEECH:
        put("src", src, torch.float32)
        put("src_length", inputs["src_length"], torch.int64)
        Ts = ((T + 1) // 2 + 1) // 2
    else:
        put("src_ids", src, torch.int64)
        put("src_padding", inputs["src_padding"], torch.float32)
        Ts = T
    bt.B, bt.T, bt.L, bt.training = B, T, 1, 0
    enc = torch.empty(B, Ts, cfg.d, dtype=torch.float32, device=rt.device)
    bias = torch.empty(B, Ts, dtype=torch.float32, device=rt.device)
    need = int(rt.lib.b200st_encode_workspace_bytes(rt.handle, B, T))
    if need <= 0:
        raise L.B200STError("workspace planning failed: " + rt.lib.b200st_last_error().decode())
    bufs = _bufs(rt, need)
    L.check(rt.lib.b200st_encode(rt.handle, C.byref(bufs), C.byref(bt), enc.data_ptr(), bias.data_ptr(), L._stream()))
    return enc, bias


def create_decoding_cache(rt, encoder_outputs, memory_bias, max_len, use_shadow=False):
    """create_decoding_internal_cache(is_inference=True) + memorize_memory for every layer."""
    B, Tm, _ = encoder_outputs.shape
    cache = DecodingCache(rt, B, Tm, max_len, memory_bias.contiguous() if memory_bias is not None else None, use_shadow)
    bufs = _bufs(rt)
    enc = encoder_outputs.to(torch.float32).contiguous()
    L.check(rt.lib.b200st_decode_init(rt.handle, C.byref(bufs), enc.data_ptr(), C.byref(cache.state), L._stream()))
    cache.memory = enc
    return cache


def decoder_step_logits(rt, symbols, cache, time):
    """symbols_to_logits_fn(symbols [B], cache, time) -> logits [B, V]."""
    if int(time) >= cache.max_len:
        raise L.B200STError("decode position %d beyond the cache length %d" % (int(time), cache.max_len))
    ids = torch.as_tensor(symbols).to(device=rt.device, dtype=torch.int64).reshape(-1).contiguous()
    cache.time.fill_(int(time))
    logits = torch.empty(cache.B, rt.config.vocab, dtype=torch.float32, device=rt.device)
    bufs = _bufs(rt)
    L.check(rt.lib.b200st_decode_step(rt.handle, C.byref(bufs), C.byref(cache.state), ids.data_ptr(), cache.time.data_ptr(),
                                      logits.data_ptr(), L._stream()))
    return logits


def greedy_search(rt, inputs, bos_id, eos_id, unk_id=None, maximum_decode_length=256, extra_decode_length=50,
                  minimum_decode_length=0, enable_unk=False, use_shadow=False, use_graph=True, cache=None, persistent=True):
    """sequence_beam_search(beam_size=1, top_k=1): returns (hypothesis int64 [B, maximum_decode_length] padded with EOS,
    log-probability [B], decoding length [B]).  `maximum_search_steps` = max(min(T' + extra, maximum), minimum)
    (beam_search.py:357-363)."""
    if cache is None:
        enc, bias = encode(rt, inputs)
        cache = create_decoding_cache(rt, enc, bias, maximum_decode_length, use_shadow)
    B, Tm = cache.B, cache.Tm
    steps = max(min(Tm + extra_decode_length, maximum_decode_length), minimum_decode_length)
    steps = min(steps, cache.max_len)
    dev = rt.device
    bos = torch.as_tensor(bos_id).to(device=dev, dtype=torch.int64).reshape(-1)
    if bos.numel() == 1:
        bos = bos.repeat(B)
    bos = bos.contiguous()
    out = torch.empty(B, steps, dtype=torch.int64, device=dev)
    length = torch.zeros(B, dtype=torch.int32, device=dev)
    logprob = torch.zeros(B, dtype=torch.float32, device=dev)
    words = torch.zeros(64, dtype=torch.int64, device=dev)
    a = L.GreedyArgs()
    a.bos_ids, a.eos_id = bos.data_ptr(), int(eos_id)
    a.unk_id = -1 if (enable_unk or unk_id is None) else int(unk_id)
    a.min_len, a.max_steps = int(minimum_decode_length), int(steps)
    a.out_ids, a.out_len, a.out_logprob = out.data_ptr(), length.data_ptr(), logprob.data_ptr()
    # persistent: the whole search is ONE cooperative kernel (grid barriers between the phases of a token); otherwise one
    # captured CUDA graph of ~40 kernels is replayed per token (use_graph) or every kernel is launched eagerly
    a.state_words, a.use_graph = words.data_ptr(), (2 if (persistent and use_graph) else int(bool(use_graph)))
    bufs = _bufs(rt)
    # the library captures the step into a CUDA graph: that needs a real (non-legacy-default) stream
    cur = torch.cuda.current_stream(dev)
    side = getattr(rt, "_decode_stream", None)
    if side is None:
        side = rt._decode_stream = torch.cuda.Stream(device=dev)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        L.check(rt.lib.b200st_greedy_search(rt.handle, C.byref(bufs), C.byref(cache.state), C.byref(a), L._stream()))
    cur.wait_stream(side)
    if use_graph and int(rt.lib.b200st_greedy_used_graph()) < 1:
        raise L.B200STError("greedy search could neither run its persistent kernel nor capture its step graph")
    if steps < maximum_decode_length:      # padded to the fixed output length with EOS (beam_search.py:428-436)
        out = torch.cat([out, torch.full((B, maximum_decode_length - steps), int(eos_id), dtype=torch.int64, device=dev)], 1)
    return out, logprob, length
