"""Reads a TensorFlow tensor-bundle checkpoint (`*.index` + `*.data-00000-of-00001`) WITHOUT TensorFlow.

The reference restores `logs-Tacotron-2/taco_pretrained/tacotron_model.ckpt-206500` through `tf.train.Saver`
(tacotron_synthesize.py:76-78, checkpoint resolved from the `checkpoint` text file at :138-139).  The bundle format
(SURVEY.md Appendix B): `.index` is an uncompressed leveldb table whose values are `BundleEntryProto`
{1: dtype, 2: TensorShapeProto{2: dim{1: size}}, 3: shard_id, 4: offset, 5: size, 6: crc32c}; tensor bytes are raw
little-endian row-major at `offset` in the data shard.
"""
from __future__ import annotations

import os
import re
import struct

import numpy as np

_MAGIC = 0xDB4775248B80FB57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64}


def _varint(buf, pos):
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _block_entries(data, off, size):
    """Yields (key, value) of one leveldb block (prefix-compressed keys, trailing restart array)."""
    blk = data[off:off + size]
    if data[off + size] != 0:
        raise ValueError('compressed leveldb block: not supported (TF writes bundles uncompressed)')
    n_restarts = struct.unpack_from('<I', blk, len(blk) - 4)[0]
    end = len(blk) - 4 - 4 * n_restarts
    pos, key = 0, b''
    while pos < end:
        shared, pos = _varint(blk, pos)
        non_shared, pos = _varint(blk, pos)
        vlen, pos = _varint(blk, pos)
        key = key[:shared] + blk[pos:pos + non_shared]
        pos += non_shared
        yield key, blk[pos:pos + vlen]
        pos += vlen


def _parse_proto(buf):
    """Minimal protobuf wire decoder -> {field: [values]} (varints as int, length-delimited as bytes, fixed32 as int)."""
    out, pos = {}, 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = struct.unpack_from('<I', buf, pos)[0]
            pos += 4
        elif wt == 1:
            v = struct.unpack_from('<Q', buf, pos)[0]
            pos += 8
        else:
            raise ValueError(f'unsupported protobuf wire type {wt}')
        out.setdefault(field, []).append(v)
    return out


def read_index(index_path):
    """-> {name: dict(dtype, shape, shard, offset, size)} for every tensor in the bundle."""
    data = open(index_path, 'rb').read()
    if struct.unpack_from('<Q', data, len(data) - 8)[0] != _MAGIC:
        raise ValueError(f'{index_path}: not a leveldb table (bad magic)')
    footer = data[-48:]
    pos = 0
    _, pos = _varint(footer, pos)          # metaindex handle
    _, pos = _varint(footer, pos)
    idx_off, pos = _varint(footer, pos)
    idx_size, pos = _varint(footer, pos)
    entries = {}
    for _, handle in _block_entries(data, idx_off, idx_size):
        boff, p = _varint(handle, 0)
        bsize, p = _varint(handle, p)
        for key, val in _block_entries(data, boff, bsize):
            if not key:
                continue                    # BundleHeaderProto
            e = _parse_proto(val)
            shape = []
            if 2 in e:
                for dim in _parse_proto(e[2][0]).get(2, []):
                    shape.append(_parse_proto(dim).get(1, [0])[0])
            entries[key.decode()] = dict(dtype=e.get(1, [0])[0], shape=tuple(shape), shard=e.get(3, [0])[0],
                                         offset=e.get(4, [0])[0], size=e.get(5, [0])[0])
    return entries


def resolve_checkpoint(path):
    """Accepts a checkpoint prefix or a directory holding TF's `checkpoint` pointer file (model_checkpoint_path: "...")."""
    if os.path.isdir(path):
        ptr = os.path.join(path, 'checkpoint')
        m = re.search(r'model_checkpoint_path:\s*"([^"]+)"', open(ptr).read())
        if not m:
            raise ValueError(f'{ptr}: no model_checkpoint_path')
        path = os.path.join(path, os.path.basename(m.group(1)))
    if not os.path.isfile(path + '.index'):
        raise FileNotFoundError(path + '.index')
    return path


def load_bundle(prefix, name_filter=None):
    """-> {name: numpy array}.  Optimizer slots (`.../Adam`, `.../Adam_1`, `*_power`) are skipped."""
    prefix = resolve_checkpoint(prefix)
    entries = read_index(prefix + '.index')
    shards = {}
    out = {}
    for name, e in entries.items():
        if name.endswith('/Adam') or name.endswith('/Adam_1') or name.endswith('_power'):
            continue
        if name_filter is not None and not name_filter(name):
            continue
        if e['dtype'] not in _DTYPES:
            continue
        shard = e['shard']
        if shard not in shards:
            num = 1
            for f in os.listdir(os.path.dirname(prefix) or '.'):
                m = re.match(re.escape(os.path.basename(prefix)) + r'\.data-\d{5}-of-(\d{5})$', f)
                if m:
                    num = int(m.group(1))
            shards[shard] = np.memmap(f'{prefix}.data-{shard:05d}-of-{num:05d}', dtype=np.uint8, mode='r')
        raw = shards[shard][e['offset']:e['offset'] + e['size']]
        arr = np.frombuffer(raw.tobytes(), dtype=_DTYPES[e['dtype']])
        out[name] = arr.reshape(e['shape']) if e['shape'] else arr.reshape(())
    return out


PREFIX = 'Tacotron_model/inference/'


def load_tacotron_weights(path):
    """Inference variables of the reference graph (tacotron/models/tacotron.py:28-152), keys without the common prefix."""
    raw = load_bundle(path, lambda n: n.startswith(PREFIX) or n == 'global_step')
    out = {k[len(PREFIX):]: v for k, v in raw.items() if k.startswith(PREFIX)}
    if 'global_step' in raw:
        out['global_step'] = raw['global_step']
    return out
