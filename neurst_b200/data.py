"""Input-pipeline edge of the speech-to-text task: frame-budget length bucketing and SpecAugment.

  create_audio_bucket_boundaries / bucket table   neurst/tasks/speech2text.py:38-56,296-384
  SpecAugment                                      neurst/utils/audio_lib.py:24-257 (time warping is a documented no-op there too)

Bucketing is host logic (which utterance goes into which padded batch); SpecAugment runs on the GPU on the padded batch
(the masks are sampled on the host exactly as `freq_or_time_masking_numpy`, the fill is one masked-select on the device),
so that augmentation does not sit on the critical path of the HostPipeline copy stream.
"""
import math

import torch


def minimal_multiple(val, factor):
    """neurst/training/training_utils.py `minimal_multiple`."""
    return int(math.ceil(val / float(factor)) * factor)


def create_audio_bucket_boundaries(maxlen, minlen=128):
    """Restates neurst/tasks/speech2text.py:38-56 (geometric-ish growth of the bucket width)."""
    if minlen is None:
        minlen = 128
    bounds = [minlen]
    base = minlen
    base_incr = int(2 ** ((math.log2(minlen) + 1) // 2))
    base_incr_mult = 1
    times = len(str(int(minlen)))
    while True:
        for _ in range(times):
            bounds.append(bounds[-1] + base)
            if bounds[-1] > maxlen:
                break
        base += base_incr * base_incr_mult
        base_incr_mult += 1
        if bounds[-1] > maxlen:
            break
    bounds[-1] = maxlen + 1
    return bounds


class FrameBudgetBucketer:
    """The (T bucket -> batch size, transcript bound) table of SpeechToText.create_and_batch_tfds (speech2text.py:296-330):
    batch size per GPU = frame budget // bucket bound, rounded up to a multiple of 8 unless `disable_batch_efficiency`;
    transcript bound from `experimental_frame_transcript_ratio`, multiple of 8.  Every replica takes the same bucket per
    step (no stragglers): `batches()` yields whole global batches of `world` x per-GPU size."""

    def __init__(self, batch_size_per_gpu, max_src_len, max_trg_len, min_src_bucket_boundary=128, world=1,
                 frame_transcript_ratio=None, disable_batch_efficiency=False):
        assert batch_size_per_gpu > max_src_len, "batch size per gpu (%d) must be greater than max_src_len (%d)" % (
            batch_size_per_gpu, max_src_len)
        self.world = world
        self.max_trg_len = minimal_multiple(max_trg_len, 8)
        b = create_audio_bucket_boundaries(max_src_len, min_src_bucket_boundary)
        b[-1] = minimal_multiple(b[-1], 8)
        self.boundaries = b
        if disable_batch_efficiency:
            self.batch_sizes = [int(batch_size_per_gpu // x) for x in b]
        else:
            self.batch_sizes = [int(minimal_multiple(batch_size_per_gpu // x, 8)) for x in b]
        if frame_transcript_ratio is None:
            self.trg_bounds = [self.max_trg_len] * len(b)
        else:
            r = frame_transcript_ratio
            tb = [int(x / (r + i * (max_src_len / self.max_trg_len - r) / len(b))) for i, x in enumerate(b)]
            self.trg_bounds = [minimal_multiple(min(t, self.max_trg_len), 8) for t in tb]

    def bucket_of(self, n_frames):
        for i, bound in enumerate(self.boundaries):
            if n_frames <= bound:
                return i
        return None         # longer than max_src_len: filtered out, as the reference's dataset filter does

    def shapes(self):
        """[(T, B per GPU, L)] — the static shape buckets (one captured CUDA graph each in the trainer)."""
        return list(zip(self.boundaries, self.batch_sizes, self.trg_bounds))

    def batches(self, examples, pad_id=0, feature_dim=80):
        """examples: iterable of dict(audio: FloatTensor [n, F], transcript: LongTensor [l]).  Yields per-replica lists of
        padded batches {src [B,T,F,1], src_length, trg [B,L], trg_length, trg_input}, drop_remainder like the reference."""
        pools = [[] for _ in self.boundaries]
        for ex in examples:
            i = self.bucket_of(ex["audio"].shape[0])
            if i is None or ex["transcript"].numel() > self.trg_bounds[i]:
                continue
            pools[i].append(ex)
            need = self.batch_sizes[i] * self.world
            if len(pools[i]) == need:
                group, pools[i] = pools[i], []
                yield [self._pad(group[r::self.world], i, pad_id, feature_dim) for r in range(self.world)]

    def _pad(self, group, i, pad_id, feature_dim):
        T, L, B = self.boundaries[i], self.trg_bounds[i], len(group)
        src = torch.zeros(B, T, feature_dim, 1)
        trg = torch.full((B, L), int(pad_id), dtype=torch.long)
        sl, tl = torch.zeros(B, dtype=torch.long), torch.zeros(B, dtype=torch.long)
        for j, ex in enumerate(group):
            n, l = ex["audio"].shape[0], ex["transcript"].numel()
            src[j, :n, :, 0] = ex["audio"]
            trg[j, :l] = ex["transcript"]
            sl[j], tl[j] = n, l
        return dict(src=src, src_length=sl, trg=trg, trg_length=tl)


class SpecAugment:
    """SpecAugment with the reference's settings table and sampling rule (audio_lib.py:27-66,107-147,231-245):
    repeat n times f ~ U[0, F), f0 ~ U[0, nu - f); positions [f0, f0+f) := mask value (utterance mean unless given); a mask
    starting at 0 is skipped (the reference's `if f0[i] == 0: continue`); time masks are capped at p * frames."""
    SETTINGS = {"LB": dict(freq_mask_n=1, freq_mask_f=27, time_mask_n=1, time_mask_t=100, time_mask_p=1.0),
                "LD": dict(freq_mask_n=2, freq_mask_f=27, time_mask_n=2, time_mask_t=100, time_mask_p=1.0),
                "SM": dict(freq_mask_n=2, freq_mask_f=15, time_mask_n=2, time_mask_t=70, time_mask_p=0.2),
                "SS": dict(freq_mask_n=2, freq_mask_f=27, time_mask_n=2, time_mask_t=70, time_mask_p=0.2)}

    def __init__(self, freq_mask_n, freq_mask_f, time_mask_n, time_mask_t, time_mask_p, mask_value=None, time_wrap_w=0):
        assert time_mask_t > 0 and freq_mask_f > 0
        self.fn, self.ff, self.tn, self.tt, self.tp, self.mask_value = freq_mask_n, freq_mask_f, time_mask_n, time_mask_t, time_mask_p, mask_value

    @classmethod
    def build(cls, setting):
        if setting is None:
            return None
        if isinstance(setting, str):
            setting = cls.SETTINGS.get(setting)
        return cls(**setting) if setting else None

    @staticmethod
    def _ranges(n, F, size, gen, p=None):
        if size < F:
            return []
        if p:
            F = min(F, math.floor(size * p))
        if F <= 0:
            return []
        out = []
        f = torch.randint(0, F, (n,), generator=gen)
        for i in range(n):
            f0 = int(torch.randint(0, size - int(f[i]), (1,), generator=gen))
            if f0 == 0:
                continue
            out.append((f0, f0 + int(f[i])))
        return out

    def masks(self, lengths, n_freq, T, generator=None):
        """Host-sampled boolean masks [B, T] (time) and [B, n_freq] (frequency) for a padded batch."""
        B = len(lengths)
        tm, fm = torch.zeros(B, T, dtype=torch.bool), torch.zeros(B, n_freq, dtype=torch.bool)
        for b in range(B):
            for a, z in self._ranges(self.fn, self.ff, n_freq, generator):
                fm[b, a:z] = True
            for a, z in self._ranges(self.tn, self.tt, int(lengths[b]), generator, self.tp):
                tm[b, a:z] = True
        return tm, fm

    def __call__(self, src, src_length, generator=None):
        """src [B,T,F,1] (any device), src_length [B] -> augmented copy (same device).  Mask value = per-utterance mean over
        the real frames (`spectrogram.mean()` of the unpadded utterance in the reference)."""
        B, T, Fd = src.shape[0], src.shape[1], src.shape[2]
        lens = src_length.to("cpu")
        tm, fm = self.masks(lens, Fd, T, generator)
        dev = src.device
        valid = (torch.arange(T, device=dev)[None, :] < src_length.to(dev)[:, None])
        if self.mask_value is None:
            tot = (src[..., 0] * valid[..., None]).sum((1, 2))
            mv = tot / (src_length.to(dev).clamp(min=1).float() * Fd)
        else:
            mv = torch.full((B,), float(self.mask_value), device=dev)
        m = (tm.to(dev)[:, :, None] | fm.to(dev)[:, None, :]) & valid[..., None]
        return torch.where(m[..., None], mv[:, None, None, None].to(src.dtype), src)
