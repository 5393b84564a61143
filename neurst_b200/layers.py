"""Host-side mirror of the reference layer API (neurst_pt/layers/**) on top of libb200st.

Same constructor arguments, call signatures, tensor shapes and error behaviour as the reference classes,
so the parity tests read like the reference's own tests; the compute is the CUDA library (no fallback).

  MultiHeadAttention / MultiHeadSelfAttention   neurst_pt/layers/attentions/multi_head_attention.py:21-263
  TransformerEncoder                            neurst_pt/layers/encoders/transformer_encoder.py:23-138
  TransformerDecoder                            neurst_pt/layers/decoders/transformer_decoder.py:23-221
"""
import ctypes as C

import torch

from neurst_b200 import lib as L
from neurst_b200.runtime import Runtime, _ptr, make_config


class _ParamView:
    """Reference-style access to one dense transform: `._kernel`, `._bias` are views into the flat arena."""

    def __init__(self, rt, kernel, bias):
        self._kernel = rt.view(kernel)
        self._bias = rt.view(bias)


class _B200Layer:
    def __init__(self, config, device="cuda"):
        self._rt = Runtime(config, device)

    @property
    def runtime(self):
        return self._rt

    def named_parameters(self):
        return self._rt.named_parameters()

    def load_parameters(self, P):
        self._rt.load_parameters(P)

    def _f32(self, t):
        return torch.as_tensor(t).to(device=self._rt.device, dtype=torch.float32).contiguous()

    def _call_with_workspace(self, fn):
        """fn(buffers_ptr, need_ptr) -> rc; first call sizes the workspace, second runs."""
        rt = self._rt
        if rt._shadow_stale:
            rt.refresh_shadow()
        need = C.c_uint64(0)
        empty = L.Buffers()
        L.check(fn(C.byref(empty), C.byref(need)))
        bufs = rt._buffers(int(need.value), False)
        L.check(fn(C.byref(bufs), None))

    def __call__(self, *a, **kw):
        return self.forward(*a, **kw)


class MultiHeadAttention(_B200Layer):
    """ Class of multi-head scaled-dot-product attention with input/output transformations. """

    _SELF = False

    def __init__(self, input_depth, num_heads, num_units, attention_key_depth=None, attention_value_depth=None,
                 output_depth=None, attention_dropout_rate=0.1, attention_type="dot_product", memory_depth=None,
                 precision="fp32", device="cuda"):
        self._input_depth = input_depth
        self._num_heads = num_heads
        self._num_units = num_units
        self._attention_key_depth = attention_key_depth or num_units
        self._attention_value_depth = attention_value_depth or num_units
        self._output_depth = output_depth or num_units
        self._attention_dropout_rate = attention_dropout_rate
        self._attention_type = attention_type
        if self._attention_key_depth % self._num_heads != 0:
            raise ValueError("query depth ({}) must be divisible by the number of "
                             "attention heads ({}).".format(self._attention_key_depth, self._num_heads))
        if self._attention_value_depth % self._num_heads != 0:
            raise ValueError("value depth ({}) must be divisible by the number of "
                             "attention heads ({}).".format(self._attention_value_depth, self._num_heads))
        if attention_type != "dot_product":
            raise NotImplementedError("att_fn for \"{}\" not implemented.".format(attention_type))
        if self._attention_key_depth != num_units or self._attention_value_depth != num_units:
            raise NotImplementedError("key/value depth different from num_units is not supported by libb200st")
        cfg = make_config(L.MODEL_MHA, num_units, num_heads, precision=precision, mha_self=self._SELF,
                          mha_din=input_depth, mha_dmem=memory_depth or input_depth, mha_dout=self._output_depth)
        super().__init__(cfg, device)
        rt = self._rt
        self._output_transform_layer = _ParamView(rt, "att.out.kernel", "att.out.bias")
        if self._SELF:
            self._qkv_transform_layer = _ParamView(rt, "att.qkv.kernel", "att.qkv.bias")
        else:
            self._q_transform_layer = _ParamView(rt, "att.q.kernel", "att.q.bias")
            self._kv_transform_layer = _ParamView(rt, "att.kv.kernel", "att.kv.bias")

    def forward(self, query, memory, memory_bias=None, cache=None, is_training=True, decode_loop_step=None):
        if is_training and self._attention_dropout_rate > 0:
            raise NotImplementedError("attention dropout of the standalone layer runs only inside the stacks/model")
        if decode_loop_step is not None:
            raise NotImplementedError("static-shape decode cache is not supported; use the concat cache")
        query = self._f32(query)
        query_is_2d = query.dim() == 2
        if query_is_2d:
            query = query.unsqueeze(1)
        if self._SELF and cache is not None:
            return self._forward_cached(query, memory_bias, cache, query_is_2d)
        memory = query if self._SELF else self._f32(memory)
        B, Tq, _ = query.shape
        Tk = memory.shape[1]
        bias = None
        if memory_bias is not None:
            bias = self._f32(memory_bias)
            if bias.dim() != 2:
                if bias.dim() != 4:
                    raise ValueError("bias tensor with {}-dim is not valid".format(bias.dim()))
                raise NotImplementedError("only [batch, length_k] biases are supported by libb200st")
        out = torch.empty(B, Tq, self._output_depth, dtype=torch.float32, device=self._rt.device)
        self._rt._shadow_stale = True   # parameters may have been written through the views
        self._call_with_workspace(lambda bufs, need: self._rt.lib.b200st_mha_forward(
            self._rt.handle, bufs, _ptr(query), None if self._SELF else _ptr(memory), _ptr(bias), B, Tq, Tk, _ptr(out),
            L._stream(), need))
        return out.squeeze(1) if query_is_2d else out

    def _forward_cached(self, query, bias, cache, query_is_2d):
        """Concat KV cache (multi_head_attention.py:226-263): project the new positions, append to cache["keys"/"values"]
        ([B, i, H, dh]) and attend over the whole cache.  Composed from library GEMM / softmax launches."""
        rt = self._rt
        H, u = self._num_heads, self._num_units
        B, Tq, din = query.shape
        P = rt.named_parameters()
        qkv = torch.empty(B * Tq, 3 * u, dtype=torch.float32, device=rt.device)
        L.gemm(query.reshape(B * Tq, din), P["att.qkv.kernel"], qkv, b_mn=True, bias=P["att.qkv.bias"].contiguous())
        q, k, v = qkv[:, :u], qkv[:, u:2 * u], qkv[:, 2 * u:]
        keys = torch.cat([self._f32(cache["keys"]).reshape(B, -1, u), k.reshape(B, Tq, u)], dim=1).contiguous()
        values = torch.cat([self._f32(cache["values"]).reshape(B, -1, u), v.reshape(B, Tq, u)], dim=1).contiguous()
        cache["keys"], cache["values"] = keys.view(B, -1, H, u // H), values.view(B, -1, H, u // H)
        Tk = keys.shape[1]
        dh = u // H
        qh = q.reshape(B, Tq, H, dh).permute(0, 2, 1, 3)
        kh = keys.view(B, Tk, H, dh).permute(0, 2, 1, 3)
        vh = values.view(B, Tk, H, dh).permute(0, 2, 1, 3)
        S = L.padded_scores(B, H, Tq, Tk, torch.float32, rt.device)
        L.gemm(qh, kh, S, alpha=dh ** -0.5)
        Pm = L.padded_scores(B, H, Tq, Tk, torch.float32, rt.device)
        L.softmax(S, Pm, bias=self._f32(bias) if bias is not None else None)
        ctx = torch.empty(B, Tq, H, dh, dtype=torch.float32, device=rt.device)
        L.gemm(Pm, vh, ctx.permute(0, 2, 1, 3), b_mn=True)
        out = torch.empty(B * Tq, self._output_depth, dtype=torch.float32, device=rt.device)
        L.gemm(ctx.reshape(B * Tq, u), P["att.out.kernel"], out, b_mn=True, bias=P["att.out.bias"].contiguous())
        out = out.view(B, Tq, -1)
        return out.squeeze(1) if query_is_2d else out


class MultiHeadSelfAttention(MultiHeadAttention):
    """ Class of multi-head scaled-dot-product self-attention with input/output transformations. """

    _SELF = True

    def forward(self, query, bias=None, cache=None, is_training=True, decode_loop_step=None):
        return super().forward(query=query, memory=query, memory_bias=bias, cache=cache, is_training=is_training,
                               decode_loop_step=decode_loop_step)


class TransformerEncoder(_B200Layer):
    """ Defines transformer encoders as described in https://arxiv.org/abs/1706.03762. """

    def __init__(self, num_layers, hidden_size, num_attention_heads, filter_size, ffn_activation="relu",
                 attention_dropout_rate=0., attention_type="dot_product", ffn_dropout_rate=0.,
                 layer_postprocess_dropout_rate=0., layer_postprocess_epsilon=1e-6, post_normalize=False,
                 return_all_layers=False, precision="fp32", device="cuda"):
        assert post_normalize or (not post_normalize and not return_all_layers), (
            "`return_all_layers` is only available when `post_normalize`=True.")
        if post_normalize:
            raise NotImplementedError("post_normalize=True is not supported by libb200st (all presets are pre-norm)")
        if ffn_activation != "relu" or attention_type != "dot_product":
            raise NotImplementedError("only relu FFN / dot_product attention are supported")
        self._params = dict(num_layers=num_layers, hidden_size=hidden_size, num_attention_heads=num_attention_heads,
                            filter_size=filter_size, ffn_activation=ffn_activation,
                            attention_dropout_rate=attention_dropout_rate, attention_type=attention_type,
                            ffn_dropout_rate=ffn_dropout_rate,
                            layer_postprocess_dropout_rate=layer_postprocess_dropout_rate,
                            layer_postprocess_epsilon=layer_postprocess_epsilon, post_normalize=post_normalize)
        cfg = make_config(L.MODEL_ENCODER, hidden_size, num_attention_heads, filter_size, enc_layers=num_layers,
                          precision=precision, ln_eps=layer_postprocess_epsilon, attention_dropout=attention_dropout_rate,
                          ffn_dropout=ffn_dropout_rate, postprocess_dropout=layer_postprocess_dropout_rate)
        super().__init__(cfg, device)
        self._seed = 0

    def forward(self, inputs, inputs_padding, is_training=True):
        x, pad = self._f32(inputs), self._f32(inputs_padding)
        B, T, d = x.shape
        out = torch.empty(B, T, d, dtype=torch.float32, device=self._rt.device)
        self._seed += 1
        if not getattr(self, "frozen_parameters", False):     # parameters may have been written through the views
            self._rt._shadow_stale = True
        self._call_with_workspace(lambda bufs, need: self._rt.lib.b200st_encoder_forward(
            self._rt.handle, bufs, _ptr(x), _ptr(pad), B, T, _ptr(out), int(is_training), self._seed, L._stream(), need))
        return out


class TransformerDecoder(_B200Layer):
    """ Defines transformer decoder as described in https://arxiv.org/abs/1706.03762. """

    def __init__(self, num_layers, hidden_size, num_attention_heads, filter_size, ffn_activation="relu",
                 attention_dropout_rate=0., attention_type="dot_product", ffn_dropout_rate=0.,
                 layer_postprocess_dropout_rate=0., layer_postprocess_epsilon=1e-6, with_encoder_decoder_attention=True,
                 post_normalize=False, precision="fp32", device="cuda"):
        if post_normalize:
            raise NotImplementedError("post_normalize=True is not supported by libb200st (all presets are pre-norm)")
        if ffn_activation != "relu" or attention_type != "dot_product":
            raise NotImplementedError("only relu FFN / dot_product attention are supported")
        self._params = dict(num_layers=num_layers, hidden_size=hidden_size, num_attention_heads=num_attention_heads,
                            filter_size=filter_size, ffn_activation=ffn_activation,
                            attention_dropout_rate=attention_dropout_rate, attention_type=attention_type,
                            ffn_dropout_rate=ffn_dropout_rate,
                            layer_postprocess_dropout_rate=layer_postprocess_dropout_rate,
                            layer_postprocess_epsilon=layer_postprocess_epsilon,
                            with_encoder_decoder_attention=with_encoder_decoder_attention, post_normalize=post_normalize)
        self._with_encoder_decoder_attention = with_encoder_decoder_attention
        cfg = make_config(L.MODEL_DECODER, hidden_size, num_attention_heads, filter_size, dec_layers=num_layers,
                          precision=precision, ln_eps=layer_postprocess_epsilon, attention_dropout=attention_dropout_rate,
                          ffn_dropout=ffn_dropout_rate, postprocess_dropout=layer_postprocess_dropout_rate,
                          with_cross_attention=with_encoder_decoder_attention)
        super().__init__(cfg, device)
        self._seed = 0

    def create_decoding_internal_cache(self, encoder_outputs, encoder_inputs_padding, is_inference=False,
                                       decode_padded_length=None):
        """ Same dictionary layout as the reference (transformer_decoder.py:119-165); `memory_bias` is kept as the
        padding-derived additive bias, plus the raw padding for the library call. """
        if is_inference:
            if decode_padded_length is not None:
                raise NotImplementedError("static-shape decode cache is not supported; use the concat cache")
            enc = self._f32(encoder_outputs)
            B = enc.shape[0]
            H = self._params["num_attention_heads"]
            dh = self._params["hidden_size"] // H
            decoding_states = {"layer_{}".format(i): {"self_attention": {
                "keys": torch.zeros(B, 0, H, dh, device=self._rt.device),
                "values": torch.zeros(B, 0, H, dh, device=self._rt.device)}}
                for i in range(self._params["num_layers"])}
        else:
            decoding_states = None
        cache = dict(decoding_states=decoding_states)
        if self._with_encoder_decoder_attention:
            pad = self._f32(encoder_inputs_padding)
            cache["memory"] = self._f32(encoder_outputs)
            cache["memory_bias"] = pad * -1.0e9          # layer_utils.input_padding_to_bias
            cache["memory_padding"] = pad
        return cache

    def forward(self, decoder_inputs, cache, is_training=True, decode_loop_step=None):
        if decode_loop_step is not None:
            raise NotImplementedError("static-shape decode cache is not supported; use the concat cache")
        x = self._f32(decoder_inputs)
        ori_ndims = x.dim()
        if ori_ndims == 2:
            x = x.unsqueeze(1)
        if cache.get("decoding_states") is not None:
            out = self._forward_cached(x, cache)
            return out.squeeze(1) if ori_ndims == 2 else out
        B, Lq, d = x.shape
        mem, pad = cache.get("memory"), cache.get("memory_padding")
        Tm = mem.shape[1] if mem is not None else 0
        out = torch.empty(B, Lq, d, dtype=torch.float32, device=self._rt.device)
        self._seed += 1
        self._rt._shadow_stale = True
        self._call_with_workspace(lambda bufs, need: self._rt.lib.b200st_decoder_forward(
            self._rt.handle, bufs, _ptr(x), _ptr(mem), _ptr(pad), B, Lq, Tm, _ptr(out), int(is_training), self._seed,
            L._stream(), need))
        return out.squeeze(1) if ori_ndims == 2 else out

    def _forward_cached(self, x, cache):
        """Incremental inference step with the concat self-attention cache (transformer_decoder.py:167-221).
        Composed from library launches (GEMM / LayerNorm / softmax); the per-layer cache tensors keep the
        reference's [B, i, H, dh] layout."""
        from neurst_b200 import layer_step
        return layer_step.decoder_step(self._rt, x, cache)
