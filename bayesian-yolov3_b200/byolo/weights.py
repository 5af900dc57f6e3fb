"""Weights of the 75 convolutions: layer table, synthetic generator, flat blob format, darknet reader.

Layer order = creation order of the reference graph = order of a darknet weight file
(/root/reference/lib_yolo/darknet.py:42-122): 52 backbone convs, then per detection scale the head convs and
the linear detection conv.  Per conv+BN layer the reference holds kernel[HWIO], gamma, beta, moving_mean,
moving_variance; per detection conv kernel[HWIO] + bias (/root/reference/lib_yolo/layers.py:545-613).

Blob ("BYW1", little endian), consumed by byolo_load_weights():
  int32 magic 0x31575942, int32 n_layers,
  n_layers x int32[6] = (k, stride, cin, cout, has_bn, dropout_flag),
  then per layer fp32: has_bn ? beta[cout], gamma[cout], mean[cout], var[cout] : bias[cout];  kernel[k,k,cin,cout].
"""
import numpy as np

MAGIC = 0x31575942
VARIANTS = ('standard', 'aleatoric', 'epistemic')


def det_channels(variant, cls_cnt):
    return 3 * (5 + cls_cnt) if variant == 'standard' else 3 * 2 * (5 + cls_cnt)


def row_width(variant, cls_cnt):
    return {'standard': 5, 'aleatoric': 14, 'epistemic': 21}[variant] + cls_cnt


def layer_table(variant, cls_cnt=2):
    """[(k, stride, cin, cout, has_bn, dropout)] x 75."""
    assert variant in VARIANTS
    t = [(3, 1, 3, 32, 1, 0)]
    c = 32
    for half, n_blocks in ((32, 1), (64, 2), (128, 8), (256, 8), (512, 4)):
        t.append((3, 2, c, 2 * half, 1, 0))
        c = 2 * half
        t += [(1, 1, c, half, 1, 0), (3, 1, half, c, 1, 0)] * n_blocks
    mc = int(variant == 'epistemic')
    for scale, (f, cin0) in enumerate(((512, 1024), (256, 768), (128, 384))):
        if scale:
            t.append((1, 1, 2 * f, f, 1, 0))
        t += [(1, 1, cin0, f, 1, mc), (3, 1, f, 2 * f, 1, mc), (1, 1, 2 * f, f, 1, mc), (3, 1, f, 2 * f, 1, mc),
              (1, 1, 2 * f, f, 1, mc), (3, 1, f, 2 * f, 1, 0), (1, 1, 2 * f, det_channels(variant, cls_cnt), 0, 0)]
    assert len(t) == 75
    return t


def synthetic(variant, cls_cnt=2, seed=0):
    """Random-init weights with O(1) activations: He-normal kernels for leaky(0.1),
    gamma~U(.8,1.2), beta~N(0,.1), mean~N(0,.1), var~U(.5,1.5) (residual-closing convs damped x0.25);
    detection kernels ~N(0, 0.04/cin), bias 0, so that logits have a std of 1-4 and few sigmoids saturate."""
    rng = np.random.default_rng(seed)
    out = []
    for li, (k, s, cin, cout, bn, _) in enumerate(layer_table(variant, cls_cnt)):
        if bn:
            std = np.sqrt(2.0 / (1.01 * k * k * cin))
            # the 3x3 that closes a residual block adds onto the trunk: damp it so 23 adds do not blow up
            damp = 0.25 if (0 < li < 52 and k == 3 and s == 1) else 1.0
            out.append(dict(kernel=(rng.standard_normal((k, k, cin, cout)) * std).astype(np.float32),
                            gamma=(rng.uniform(0.8, 1.2, cout) * damp).astype(np.float32),
                            beta=(rng.standard_normal(cout) * 0.1 * damp).astype(np.float32),
                            mean=(rng.standard_normal(cout) * 0.1).astype(np.float32),
                            var=rng.uniform(0.5, 1.5, cout).astype(np.float32)))
        else:
            out.append(dict(kernel=(rng.standard_normal((k, k, cin, cout)) * (0.2 / np.sqrt(cin))).astype(np.float32),
                            bias=np.zeros(cout, np.float32)))
    return out


def pack(variant, weights, cls_cnt=2):
    table = layer_table(variant, cls_cnt)
    assert len(weights) == len(table)
    parts = [np.array([MAGIC, len(table)], np.int32).tobytes(), np.array(table, np.int32).tobytes()]
    for (k, s, cin, cout, bn, _), w in zip(table, weights):
        kern = np.ascontiguousarray(w['kernel'], np.float32)
        assert kern.shape == (k, k, cin, cout), (kern.shape, (k, k, cin, cout))
        names = ('beta', 'gamma', 'mean', 'var') if bn else ('bias',)
        for n in names:
            v = np.ascontiguousarray(w[n], np.float32)
            assert v.shape == (cout,)
            parts.append(v.tobytes())
        parts.append(kern.tobytes())
    return b''.join(parts)


def unpack(blob):
    hdr = np.frombuffer(blob, np.int32, 2)
    assert int(hdr[0]) == MAGIC, 'not a BYW1 blob'
    n = int(hdr[1])
    table = np.frombuffer(blob, np.int32, 6 * n, 8).reshape(n, 6)
    off = 8 + 24 * n
    out = []
    for k, s, cin, cout, bn, _ in table:
        w = {}
        for name in (('beta', 'gamma', 'mean', 'var') if bn else ('bias',)):
            w[name] = np.frombuffer(blob, np.float32, cout, off).copy()
            off += 4 * cout
        cnt = k * k * cin * cout
        w['kernel'] = np.frombuffer(blob, np.float32, cnt, off).reshape(k, k, cin, cout).copy()
        off += 4 * cnt
        out.append(w)
    assert off == len(blob)
    return out


def read_darknet(path, table, into=None):
    """Read a darknet weight file (e.g. darknet53.conv.74) covering the first len(...) layers of `table`.
    File layout per /root/reference/lib_yolo/darknet.py:42-66: 5 x int32 header, then per conv
    [beta, gamma, mean, var] (BN layers) or [bias], then the kernel as [n, c, h, w].  Returns the list of
    weight dicts for the layers the file covers; the file must be consumed exactly (darknet.py:66)."""
    with open(path, 'rb') as f:
        np.fromfile(f, dtype=np.int32, count=5)
        flat = np.fromfile(f, dtype=np.float32)
    ptr = 0
    out = [] if into is None else into
    for li, (k, s, cin, cout, bn, _) in enumerate(table):
        if ptr == len(flat):
            break
        w = {}
        for name in (('beta', 'gamma', 'mean', 'var') if bn else ('bias',)):
            w[name] = flat[ptr:ptr + cout].copy()
            ptr += cout
        cnt = k * k * cin * cout
        assert ptr + cnt <= len(flat), 'darknet weight file truncated'
        w['kernel'] = np.transpose(flat[ptr:ptr + cnt].reshape(cout, cin, k, k), (2, 3, 1, 0)).copy()
        ptr += cnt
        if into is None:
            out.append(w)
        else:
            out[li] = w
    assert ptr == len(flat), 'darknet weight file not fully consumed'
    return out


def write_darknet(path, weights, table):
    """Inverse of read_darknet (used by tests to fabricate a .conv.74-style file)."""
    with open(path, 'wb') as f:
        np.zeros(5, np.int32).tofile(f)
        for (k, s, cin, cout, bn, _), w in zip(table, weights):
            for name in (('beta', 'gamma', 'mean', 'var') if bn else ('bias',)):
                np.asarray(w[name], np.float32).tofile(f)
            np.transpose(np.asarray(w['kernel'], np.float32), (3, 2, 0, 1)).tofile(f)
