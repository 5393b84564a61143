"""TFRecord files of tf.train.Example — the on-disk format the reference trains the SpeechTransformer from — without
TensorFlow ("to feed real MuST-C TFRecords").

Host mirror of
  load_tfrecords / glob_tfrecords / parse_tfexample / take_one_record   neurst/data/dataset_utils.py:240-325,550-566
  AudioTFRecordDataset (fields, status, build_iterator)                 neurst/data/datasets/audio/audio_dataset.py:249-365
  the writer used by the reference's dataset converters                 neurst/cli/create_tfrecords.py (tf.io.TFRecordWriter)

Record framing and CRC-32C are C (libb200st_io.so, include/b200st_io.h); the Example protobuf
(tensorflow/core/example/{example,feature}.proto — third-party, restated from the published schema:
Example{1: Features{1: map<string, Feature{oneof 1: BytesList, 2: FloatList, 3: Int64List; each {repeated 1: value}}>}})
is decoded here: packed float lists become zero-copy numpy views of the file image.  Pinned on the reference's own
fixtures tests/examples/train.tfrecords-0000?-of-00004 (tests/golden/tfrecord_fixture.json).
"""
import ctypes as C
import glob as _glob
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class TFRecordError(IOError):
    """Damaged / truncated record (tf.errors.DataLossError in the reference's pipeline)."""


def io_lib():
    """ctypes handle of libb200st_io.so (built in-tree with gcc on first use)."""
    global _LIB
    if _LIB is None:
        from neurst_b200.csrc import build as _build
        path = _build.build_io()
        lib = C.CDLL(path)
        lib.b200st_crc32c.restype = C.c_uint32
        lib.b200st_crc32c.argtypes = [C.c_uint32, C.c_void_p, C.c_size_t]
        lib.b200st_crc32c_mask.restype = C.c_uint32
        lib.b200st_crc32c_mask.argtypes = [C.c_uint32]
        lib.b200st_tfrecord_index.restype = C.c_int64
        lib.b200st_tfrecord_index.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
        lib.b200st_tfrecord_frame.restype = None
        lib.b200st_tfrecord_frame.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p]
        lib.b200st_crc32c_table.restype = C.c_uint32
        lib.b200st_crc32c_table.argtypes = [C.c_uint32, C.c_void_p, C.c_size_t]
        lib.b200st_example_lookup.restype = C.c_int
        lib.b200st_example_lookup.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_char_p), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.b200st_decode_varints.restype = C.c_int64
        lib.b200st_decode_varints.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int64]
        lib.b200st_pad_rows_f32.restype = None
        lib.b200st_pad_rows_f32.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32]
        if lib.b200st_io_version() != 2:
            raise RuntimeError("libb200st_io.so does not match this binding")
        _LIB = lib
    return _LIB


def crc32c(data, crc=0):
    buf = bytes(data) if not isinstance(data, (bytes, bytearray)) else data
    return int(io_lib().b200st_crc32c(crc, C.c_char_p(bytes(buf)), len(buf)))


def masked_crc32c(data):
    return int(io_lib().b200st_crc32c_mask(crc32c(data)))


# ------------------------------------------------------------------------------------------------ record framing
def read_records(path, verify=2):
    """Yields the payload of every record of one file as a uint8 numpy view of the file image (one read of the whole file,
    like the reference's TFRecordDataset(buffer_size=128 MB)).  verify: 0 none / 1 length CRCs / 2 length + payload CRCs."""
    if os.path.getsize(path) == 0:
        return
    img = np.memmap(path, dtype=np.uint8, mode="r")      # pages stream in as the index / checksum pass walks the file
    lib = io_lib()
    n = int(lib.b200st_tfrecord_index(img.ctypes.data, img.size, None, None, 0, int(verify)))
    if n < 0:
        raise TFRecordError("%s: corrupted or truncated record at byte %d" % (path, -1 - n))
    off = np.empty(n, np.int64)
    ln = np.empty(n, np.int64)
    lib.b200st_tfrecord_index(img.ctypes.data, img.size, off.ctypes.data, ln.ctypes.data, n, 0)
    for i in range(n):
        yield img[off[i]:off[i] + ln[i]]


class TFRecordWriter:
    """tf.io.TFRecordWriter (no compression)."""

    def __init__(self, path):
        self._f = open(path, "wb")

    def write(self, payload):
        payload = bytes(payload)
        head, foot = (C.c_uint8 * 12)(), (C.c_uint8 * 4)()
        io_lib().b200st_tfrecord_frame(payload, len(payload), head, foot)
        self._f.write(bytes(head))
        self._f.write(payload)
        self._f.write(bytes(foot))

    def close(self):
        if self._f:
            self._f.close()
            self._f = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


# ------------------------------------------------------------------------------------------------ protobuf wire format
def _varint(buf, pos):
    x = shift = 0
    while True:
        b = int(buf[pos]); pos += 1
        x |= (b & 0x7f) << shift
        if b < 0x80:
            return x, pos
        shift += 7
        if shift > 63:
            raise TFRecordError("malformed varint")


def _fields(buf, pos, end):
    """(field number, wire type, value | (start, stop)) of one message."""
    while pos < end:
        key, pos = _varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
            yield num, wt, v
        elif wt == 2:
            n, pos = _varint(buf, pos)
            if pos + n > end:
                raise TFRecordError("length-delimited field runs past its message")
            yield num, wt, (pos, pos + n)
            pos += n
        elif wt == 5:
            yield num, wt, (pos, pos + 4)
            pos += 4
        elif wt == 1:
            yield num, wt, (pos, pos + 8)
            pos += 8
        else:
            raise TFRecordError("unsupported wire type %d" % wt)
    if pos != end:
        raise TFRecordError("message overruns its length")


def _signed64(x):
    return x - (1 << 64) if x >= (1 << 63) else x


def _packed_varints(buf, a, b):
    seg = np.asarray(buf[a:b])
    if seg.size == 0:
        return np.empty(0, np.int64)
    if (seg < 0x80).all():                      # every value < 128 (one byte each): the common case for short id lists
        return seg.astype(np.int64)
    ends = np.flatnonzero(seg < 0x80)
    if ends.size == 0 or ends[-1] != seg.size - 1:
        raise TFRecordError("malformed packed varint list")
    starts = np.concatenate(([0], ends[:-1] + 1))
    out = np.zeros(ends.size, np.uint64)
    width = ends - starts + 1
    for k in range(int(width.max())):           # vectorised over the values, one pass per byte position
        sel = width > k
        out[sel] |= (seg[starts[sel] + k].astype(np.uint64) & np.uint64(0x7f)) << np.uint64(min(7 * k, 63))
    return out.view(np.int64)


def _feature(buf, a, b):
    """Feature -> ('bytes', [bytes]) | ('float', float32 array) | ('int64', int64 array)."""
    kind, val = None, None
    for num, wt, v in _fields(buf, a, b):
        if wt != 2 or num not in (1, 2, 3):
            continue
        la, lb = v
        if num == 1:
            kind = "bytes"
            val = [bytes(buf[x:y]) for n2, w2, (x, y) in ((n_, w_, v_) for n_, w_, v_ in _fields(buf, la, lb) if w_ == 2 and n_ == 1)]
        elif num == 2:
            kind = "float"
            parts = []
            for n2, w2, v2 in _fields(buf, la, lb):
                if n2 != 1:
                    continue
                if w2 == 2:                      # packed
                    parts.append(np.frombuffer(buf[v2[0]:v2[1]], dtype="<f4"))
                elif w2 == 5:                    # one unpacked value
                    parts.append(np.frombuffer(buf[v2[0]:v2[1]], dtype="<f4"))
            val = parts[0] if len(parts) == 1 else (np.concatenate(parts) if parts else np.empty(0, np.float32))
        else:
            kind = "int64"
            parts = []
            for n2, w2, v2 in _fields(buf, la, lb):
                if n2 != 1:
                    continue
                if w2 == 2:
                    parts.append(_packed_varints(buf, v2[0], v2[1]))
                elif w2 == 0:
                    parts.append(np.array([_signed64(v2)], np.int64))
            val = parts[0] if len(parts) == 1 else (np.concatenate(parts) if parts else np.empty(0, np.int64))
    return kind, val


class FeatureLookup:
    """Native lookup of a fixed set of features in serialized Examples (b200st_example_lookup): one C call per record instead
    of a Python walk over the protobuf fields; float lists come back as zero-copy views, id lists are decoded in C."""

    def __init__(self, name_to_features):
        self.names = list(name_to_features)
        self.kinds = [name_to_features[k] for k in self.names]
        n = len(self.names)
        self._keys = (C.c_char_p * n)(*[k.encode("utf-8") for k in self.names])
        self._kind = np.zeros(n, np.int32)
        self._off, self._len, self._cnt = np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n, np.int64)
        self._lib = io_lib()

    _CODE = {1: "bytes", 2: "float", 3: "int64"}

    def __call__(self, rec):
        """rec: uint8 numpy view of one record -> {name: value} like `to_dense(parse_example(rec), schema)`; None when the
        record stores a list unpacked (the caller falls back to the general decoder)."""
        rc = self._lib.b200st_example_lookup(rec.ctypes.data, rec.size, self._keys, len(self.names), self._kind.ctypes.data,
                                             self._off.ctypes.data, self._len.ctypes.data, self._cnt.ctypes.data)
        if rc != 0:
            raise TFRecordError("malformed tf.train.Example")
        out = {}
        for i, name in enumerate(self.names):
            want, got = self.kinds[i], int(self._kind[i])
            a, ln, cnt = int(self._off[i]), int(self._len[i]), int(self._cnt[i])
            if got == 0:
                out[name] = [] if want == "bytes" else np.empty(0, np.float32 if want == "float" else np.int64)
                continue
            if self._CODE[got] != want:
                raise TFRecordError("feature %r is stored as %s_list, the schema asks for %s" % (name, self._CODE[got], want))
            if cnt < 0 or (got == 1 and cnt != 1):
                return None
            if got == 2:
                out[name] = np.frombuffer(rec[a:a + ln], dtype="<f4")
            elif got == 3:
                v = np.empty(cnt, np.int64)
                if self._lib.b200st_decode_varints(rec.ctypes.data + a, ln, v.ctypes.data, cnt) != cnt:
                    raise TFRecordError("malformed packed varint list")
                out[name] = v
            else:
                out[name] = [bytes(rec[a:a + ln])]
        return out


def parse_example(payload):
    """Serialized tf.train.Example -> {name: (kind, value)} with kind in {'bytes','float','int64', None (empty feature)}."""
    buf = payload if isinstance(payload, np.ndarray) else np.frombuffer(bytes(payload), np.uint8)
    out = {}
    for num, wt, v in _fields(buf, 0, buf.size):
        if num != 1 or wt != 2:
            continue
        for n2, w2, v2 in _fields(buf, v[0], v[1]):            # Features.feature map entries
            if n2 != 1 or w2 != 2:
                continue
            key, feat = None, (None, None)
            for n3, w3, v3 in _fields(buf, v2[0], v2[1]):
                if n3 == 1 and w3 == 2:
                    key = bytes(buf[v3[0]:v3[1]]).decode("utf-8")
                elif n3 == 2 and w3 == 2:
                    feat = _feature(buf, v3[0], v3[1])
            if key is not None:
                out[key] = feat
    return out


def _enc_varint(x):
    x &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = x & 0x7f
        x >>= 7
        if x:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _ld(num, body):
    return _enc_varint((num << 3) | 2) + _enc_varint(len(body)) + body


def encode_example(features):
    """{name: float array | int array | bytes | str | list of bytes/str} -> serialized tf.train.Example (packed lists, map
    entries in sorted key order — what the protobuf C++ serializer TF uses emits with deterministic serialization)."""
    entries = []
    for key in sorted(features):
        v = features[key]
        if isinstance(v, (bytes, str)):
            v = [v]
        if isinstance(v, (list, tuple)) and (len(v) == 0 or isinstance(v[0], (bytes, str))):
            body = b"".join(_ld(1, x.encode("utf-8") if isinstance(x, str) else x) for x in v)
            feat = _ld(1, body)
        else:
            a = np.asarray(v)
            if a.dtype.kind == "f":
                feat = _ld(2, _ld(1, a.astype("<f4").tobytes()) if a.size else b"")
            elif a.dtype.kind in "iub":
                feat = _ld(3, _ld(1, b"".join(_enc_varint(int(x)) for x in a.reshape(-1))) if a.size else b"")
            else:
                raise TypeError("feature %r: unsupported dtype %s" % (key, a.dtype))
        entries.append(_ld(1, _ld(1, key.encode("utf-8")) + _ld(2, feat)))
    return _ld(1, b"".join(entries))


# ------------------------------------------------------------------------------------------------ dataset_utils mirror
def _flatten(x):
    if isinstance(x, (list, tuple)):
        return [z for y in x for z in _flatten(y)]
    return [s for s in str(x).split(",") if s]


def glob_tfrecords(file_path):
    """dataset_utils.glob_tfrecords + Dataset.list_files(shuffle=False): directory -> '*train*', existing file, else
    prefix + '*'; the matched names in sorted order."""
    files = []
    for f in _flatten(file_path):
        if os.path.isdir(f):
            files.extend(_glob.glob(os.path.join(f, "*train*")))
        elif os.path.exists(f):
            files.append(f)
        else:
            files.extend(_glob.glob(f + "*"))
    return sorted(files)


VarLenFloat, VarLenInt64, VarLenString = "float", "int64", "bytes"


def to_dense(parsed, name_to_features, feature_name_mapping=None):
    """parse_single_example(VarLenFeature) + tf.sparse.to_dense: a missing / empty feature is an empty array of the declared
    type; a feature stored with another type raises (TF: 'Data types don't match')."""
    out = {}
    for name, kind in name_to_features.items():
        got_kind, val = parsed.get(name, (None, None))
        if got_kind is None:
            val = [] if kind == "bytes" else np.empty(0, np.float32 if kind == "float" else np.int64)
        elif got_kind != kind:
            raise TFRecordError("feature %r is stored as %s_list, the schema asks for %s" % (name, got_kind, kind))
        out[(feature_name_mapping or {}).get(name, name)] = val
    return out


def load_tfrecords(file_path, name_to_features=None, feature_name_mapping=None, map_func=None, sharding_index=0,
                   num_shards=1, auxiliary_elements=None, cycle_length=10, verify=2):
    """Generator with the deterministic order of dataset_utils.load_tfrecords(shuffle=False): sorted file list, shard
    `sharding_index` of `num_shards` by FILE (Dataset.shard on the file names), then interleave(cycle_length=10,
    block_length=1): one record from each of the (up to) 10 open files in turn; an exhausted file is replaced by the next
    unopened one in its slot."""
    files = glob_tfrecords(file_path)
    if num_shards > 1:
        files = files[sharding_index::num_shards]
    lookup = FeatureLookup(name_to_features) if name_to_features is not None else None
    pending = iter(files)
    slots = []
    for _ in range(cycle_length):
        f = next(pending, None)
        if f is None:
            break
        slots.append(read_records(f, verify))
    i = 0
    while slots:
        i %= len(slots)
        rec = next(slots[i], None)
        if rec is None:
            f = next(pending, None)
            if f is None:
                slots.pop(i)            # the following slots move up: the cycle goes on with the next file
            else:
                slots[i] = read_records(f, verify)
            continue
        i += 1
        if name_to_features is None:
            yield rec
            continue
        fast = lookup(rec)
        if fast is None:           # a list stored unpacked / several bytes values: general decoder
            el = to_dense(parse_example(rec), name_to_features, feature_name_mapping)
        else:
            el = {(feature_name_mapping or {}).get(k, k): v for k, v in fast.items()}
        if isinstance(auxiliary_elements, dict):
            el.update(auxiliary_elements)
        yield el if map_func is None else map_func(el)


def take_one_record(data_path):
    """dataset_utils.take_one_record: the parsed first Example of the first file."""
    for rec in load_tfrecords(_flatten(data_path)[0]):
        return parse_example(rec)
    raise TFRecordError("no record under %r" % (data_path,))


def _python_type(v):
    """to_numpy_or_python_type(bytes_as_str=True) (neurst/utils/misc.py:86-126): a bytes list becomes its FIRST element as
    str; numeric arrays stay arrays."""
    if isinstance(v, list):
        return v[0].decode("utf-8") if v else np.empty(0, object)
    return v


class AudioTFRecordDataset:
    """`audio_tfrecord` dataset: records {audio: float_list (extracted features, flattened [frames*dim]) | int64_list (raw
    samples), transcript: bytes | int64_list (already tokenised), src_lang, uuid}."""

    def __init__(self, args):
        self._data_path = args["data_path"]
        self._feature_key = args.get("feature_key", "audio")
        self._transcript_key = args.get("transcript_key", "transcript")
        ex = take_one_record(self._data_path)
        kind = ex.get(self._feature_key, (None, None))[0]
        if kind == "float":
            self._audio_is_extracted = True
        elif kind == "int64":
            self._audio_is_extracted = False
        else:
            raise ValueError("Fail to read %s" % (self._data_path,))
        self._transcript_is_projected = ex.get(self._transcript_key, (None, None))[0] == "int64"

    @property
    def status(self):
        return {"audio": "projected" if self._audio_is_extracted else "raw",
                "transcript": "projected" if self._transcript_is_projected else "raw"}

    @property
    def fields(self):
        return {self._feature_key: VarLenFloat if self._audio_is_extracted else VarLenInt64,
                self._transcript_key: VarLenInt64 if self._transcript_is_projected else VarLenString,
                "src_lang": VarLenString, "uuid": VarLenString}

    def build_iterator(self, map_func=None, shard_id=0, total_shards=1):
        def gen():
            for x in load_tfrecords(self._data_path, name_to_features=self.fields, sharding_index=shard_id,
                                    num_shards=total_shards,
                                    feature_name_mapping={self._feature_key: "audio", self._transcript_key: "transcript"}):
                data = {k: _python_type(v) for k, v in x.items()}
                yield data if map_func is None else map_func(data)
        return gen
