"""Weights from a TensorFlow-1.x checkpoint of the reference models, without TensorFlow.

Replaces `tf.train.Saver().restore(sess, checkpoint)` of the reference scripts (/root/reference/detect.py:96-107,
inference_epistemic.py:27-36, lib_yolo/train.py:97-102) as the way trained ECP models enter the engine:

    vars = read_bundle('/path/model.ckpt-500000')                 # {variable name: ndarray}
    weights = weights_from_variables('epistemic', 2, vars)        # byolo.weights layer list -> Engine.load_weights

Two parts:
* `variable_names` / `weights_from_variables`: the variable NAMES.  The scope structure comes from the reference's own
  graph code (`variable_scope(None, default_name='conv' | 'downsample')` inside 'darknet53' / 'det_net_j', the
  'detection' scope, model.py:43-66, :110, yolov3.py:240-284) with TensorFlow's rule for making default names unique
  (conv, conv_1, ...).  Pinned by tests/golden/tf_variable_names.json, which tests/golden/gen_golden.py records while
  it executes that code on the numpy TF stand-in.
* `read_bundle`: the tensor-bundle file format (`<prefix>.index` = a LevelDB-style sorted table of BundleEntryProto,
  `<prefix>.data-0000k-of-0000n` = raw little-endian tensors).  TensorFlow is not installable here, so this reader is
  a restatement of the published format checked only against the writer in tests/test_host_logic.py (same
  restatement): FORMAT PARITY UNPINNED until it has read a checkpoint written by TensorFlow itself.
"""
import os
import struct

import numpy as np

from . import weights as _weights

_BN_LEAVES = ('gamma', 'beta', 'moving_mean', 'moving_variance')      # creation order inside tf.layers.batch_normalization
_BN_KEYS = ('gamma', 'beta', 'mean', 'var')                           # the same four in byolo.weights dicts


def _scopes(variant, cls_cnt=2):
    """Scope of each of the 75 convs in creation order: 'darknet53/conv_3', 'det_net_2/detection', ..."""
    table = _weights.layer_table(variant, cls_cnt)
    out, counters = [], {}

    def unique(parent, default):
        n = counters.get((parent, default), 0)
        counters[(parent, default)] = n + 1
        return '%s/%s' % (parent, default if n == 0 else '%s_%d' % (default, n))

    for li, (k, s, cin, cout, bn, _) in enumerate(table):
        if li < 52:
            parent = 'darknet53'
        else:
            parent = 'det_net_%d' % (1 + (li >= 59) + (li >= 67))          # 7 convs in head 1, 8 in heads 2 and 3
        out.append('%s/detection' % parent if not bn else unique(parent, 'downsample' if s == 2 else 'conv'))
    return out


def variable_names(variant, cls_cnt=2):
    """[(name, shape)] of all checkpoint variables of the inference graph, in creation order."""
    out = []
    for scope, (k, s, cin, cout, bn, _) in zip(_scopes(variant, cls_cnt), _weights.layer_table(variant, cls_cnt)):
        out.append((scope + '/conv2d/kernel', (k, k, cin, cout)))
        if bn:
            out += [('%s/batch_normalization/%s' % (scope, leaf), (cout,)) for leaf in _BN_LEAVES]
        else:
            out.append((scope + '/conv2d/bias', (cout,)))
    return out


def weights_from_variables(variant, cls_cnt, variables):
    """{variable name: array} (names with or without the ':0' suffix; optimizer slots etc. are ignored) -> the 75
    weight dicts of byolo.weights.  Raises KeyError / ValueError on a missing variable or a wrong shape."""
    def get(name, shape):
        v = variables[name] if name in variables else variables[name + ':0']
        v = np.asarray(v, np.float32)
        if tuple(v.shape) != tuple(shape):
            raise ValueError('%s has shape %s, expected %s' % (name, v.shape, shape))
        return v

    out = []
    for scope, (k, s, cin, cout, bn, _) in zip(_scopes(variant, cls_cnt), _weights.layer_table(variant, cls_cnt)):
        w = {'kernel': get(scope + '/conv2d/kernel', (k, k, cin, cout))}           # HWIO, as tf.layers.conv2d stores it
        if bn:
            for leaf, key in zip(_BN_LEAVES, _BN_KEYS):
                w[key] = get('%s/batch_normalization/%s' % (scope, leaf), (cout,))
        else:
            w['bias'] = get(scope + '/conv2d/bias', (cout,))
        out.append(w)
    return out


def variables_from_weights(variant, cls_cnt, weights):
    """Inverse of weights_from_variables (tests, and exporting engine weights under the reference's names)."""
    out = {}
    for scope, w in zip(_scopes(variant, cls_cnt), weights):
        out[scope + '/conv2d/kernel'] = np.asarray(w['kernel'], np.float32)
        if 'bias' in w:
            out[scope + '/conv2d/bias'] = np.asarray(w['bias'], np.float32)
        else:
            for leaf, key in zip(_BN_LEAVES, _BN_KEYS):
                out['%s/batch_normalization/%s' % (scope, leaf)] = np.asarray(w[key], np.float32)
    return out


# ----------------------------------------------------------------------------------------------------------------------
# tensor bundle reader
# ----------------------------------------------------------------------------------------------------------------------
_TABLE_MAGIC = 0xdb4775248b80fb57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 19: np.float16}      # tensorflow DataType enum


def _varint(buf, pos):
    result = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _snappy_decompress(src):
    """Raw snappy block format (tables may be written with snappy compression)."""
    n, pos = _varint(src, 0)
    out = bytearray()
    while pos < len(src):
        tag = src[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:                                    # literal
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(src[pos:pos + nb], 'little')
                pos += nb
            ln += 1
            out += src[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln, off = ((tag >> 2) & 7) + 4, ((tag >> 5) << 8) | src[pos]
            pos += 1
        elif kind == 2:
            ln, off = (tag >> 2) + 1, int.from_bytes(src[pos:pos + 2], 'little')
            pos += 2
        else:
            ln, off = (tag >> 2) + 1, int.from_bytes(src[pos:pos + 4], 'little')
            pos += 4
        if off == 0 or off > len(out):
            raise ValueError('corrupt snappy stream')
        if off >= ln:                                    # plain back-reference: one slice
            start = len(out) - off
            out += out[start:start + ln]
        else:                                            # overlapping copy = the last `off` bytes repeated
            pat = bytes(out[-off:])
            out += (pat * (ln // off + 1))[:ln]
    if len(out) != n:
        raise ValueError('corrupt snappy stream (length)')
    return bytes(out)


def _read_block(data, offset, size):
    """Block contents at a BlockHandle: `size` bytes followed by a 1-byte compression type and a 4-byte crc."""
    raw, ctype = data[offset:offset + size], data[offset + size]
    if ctype == 1:
        raw = _snappy_decompress(raw)
    elif ctype != 0:
        raise ValueError('unsupported table block compression %d' % ctype)
    return raw


def _block_entries(block):
    """(key, value) pairs of one table block: prefix-compressed entries, then the restart array."""
    n_restarts = struct.unpack_from('<I', block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos, key = 0, b''
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def _proto_fields(buf):
    """Minimal protobuf wire reader: yields (field number, wire type, value)."""
    pos = 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v, pos = buf[pos:pos + 8], pos + 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v, pos = buf[pos:pos + ln], pos + ln
        elif wt == 5:
            v, pos = buf[pos:pos + 4], pos + 4
        else:
            raise ValueError('unsupported protobuf wire type %d' % wt)
        yield field, wt, v


def _parse_entry(buf):
    """BundleEntryProto: dtype = 1, shape = 2 (TensorShapeProto: repeated dim = 2 {size = 1}), shard_id = 3, offset = 4,
    size = 5, crc32c = 6, slices = 7."""
    e = dict(dtype=0, shape=[], shard_id=0, offset=0, size=0, sliced=False)
    for field, wt, v in _proto_fields(buf):
        if field == 1:
            e['dtype'] = v
        elif field == 2:
            for f2, _, dim in _proto_fields(v):
                if f2 == 2:
                    size = 0
                    for f3, _, x in _proto_fields(dim):
                        if f3 == 1:
                            size = x
                    e['shape'].append(size)
        elif field == 3:
            e['shard_id'] = v
        elif field == 4:
            e['offset'] = v
        elif field == 5:
            e['size'] = v
        elif field == 7:
            e['sliced'] = True
    return e


def read_bundle(prefix, names=None):
    """{name: ndarray} for the tensors of a checkpoint `prefix` (the path without .index / .data-...).  `names`
    restricts the result (e.g. to the inference variables); partitioned (sliced) variables are not supported."""
    with open(prefix + '.index', 'rb') as f:
        idx = f.read()
    if len(idx) < 48 or struct.unpack_from('<Q', idx, len(idx) - 8)[0] != _TABLE_MAGIC:
        raise ValueError('%s.index is not a tensor-bundle index (bad table magic)' % prefix)
    footer = idx[len(idx) - 48:]
    _, pos = _varint(footer, 0)                          # metaindex handle (offset, size): unused
    _, pos = _varint(footer, pos)
    ioff, pos = _varint(footer, pos)                     # index block handle
    isize, pos = _varint(footer, pos)
    entries, num_shards = {}, 1
    for _, handle in _block_entries(_read_block(idx, ioff, isize)):
        boff, p2 = _varint(handle, 0)
        bsize, _ = _varint(handle, p2)
        for key, value in _block_entries(_read_block(idx, boff, bsize)):
            if key == b'':                               # BundleHeaderProto: num_shards = 1, endianness = 2
                for field, _, v in _proto_fields(value):
                    if field == 1:
                        num_shards = v
                    elif field == 2 and v != 0:
                        raise ValueError('big-endian checkpoints are not supported')
            else:
                entries[key.decode()] = _parse_entry(value)
    shards, out = {}, {}
    for name, e in entries.items():
        if names is not None and name not in names:
            continue
        if e['sliced']:
            raise ValueError('%s is a partitioned variable (not supported)' % name)
        if e['dtype'] not in _DTYPES:
            continue                                     # strings, resources ...: nothing the model needs
        sid = e['shard_id']
        if sid not in shards:
            shards[sid] = np.memmap('%s.data-%05d-of-%05d' % (prefix, sid, num_shards), dtype=np.uint8, mode='r')
        dt = np.dtype(_DTYPES[e['dtype']])
        cnt = int(np.prod(e['shape'])) if e['shape'] else 1
        if cnt * dt.itemsize != e['size']:
            raise ValueError('%s: %d bytes stored for shape %s' % (name, e['size'], e['shape']))
        out[name] = np.frombuffer(shards[sid][e['offset']:e['offset'] + e['size']].tobytes(), dt).reshape(e['shape'])
    return out


def load(prefix, variant, cls_cnt=2):
    """checkpoint prefix -> byolo.weights layer list (what Engine.load_weights takes)."""
    wanted = {n for n, _ in variable_names(variant, cls_cnt)}
    variables = read_bundle(prefix, wanted)
    missing = sorted(n for n in wanted if n not in variables and n + ':0' not in variables)
    if missing:          # a checkpoint of another model class / cls_cnt must not fall through to a half-initialised engine
        raise KeyError('checkpoint %s lacks %d of the %d variables of the %s model (cls_cnt %d): %s%s' % (
            prefix, len(missing), len(wanted), variant, cls_cnt, ', '.join(missing[:8]), ' ...' if len(missing) > 8 else ''))
    return weights_from_variables(variant, cls_cnt, variables)


def latest_checkpoint(directory):
    """tf.train.latest_checkpoint: the prefix named by the `checkpoint` state file of a directory, or None."""
    state = os.path.join(directory, 'checkpoint')
    if not os.path.exists(state):
        return None
    with open(state) as f:
        for line in f:
            if line.startswith('model_checkpoint_path:'):
                p = line.split(':', 1)[1].strip().strip('"')
                return p if os.path.isabs(p) else os.path.join(directory, p)
    return None
