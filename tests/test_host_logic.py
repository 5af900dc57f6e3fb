"""CPU tests (no GPU): C-ABI surface, weight formats, reference-surface shims, ECP records, sharding + gather (gloo)."""
import ctypes
import json
import os
import re
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from byolo import _lib
    hdr = open(os.path.join(ROOT, 'include', 'byolo.h')).read()
    declared = set(re.findall(r'\b(byolo_[a-z0-9_]+)\s*\(', hdr))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.lib().byolo_version() == 2


def test_no_cpu_fallback_and_argument_errors():
    from byolo import _lib
    lib = _lib.lib()
    h = ctypes.c_void_p()
    cfg = _lib.Config(variant=2, height=100, width=160, cls_cnt=2, max_batch=1, T=4, precision=2, drop_prob=0.1)
    assert lib.byolo_create(ctypes.byref(cfg), ctypes.byref(h)) == -1          # 100 % 32 != 0 (yolov3.py:207-211)
    assert b'multiple of 32' in lib.byolo_last_error()
    if not torch.cuda.is_available():
        cfg.height = 96
        assert lib.byolo_create(ctypes.byref(cfg), ctypes.byref(h)) == -2
        assert b'no CPU fallback' in lib.byolo_last_error()
        import byolo
        with pytest.raises(RuntimeError):
            byolo.Engine('standard', (96, 160))


def test_weight_blob_and_darknet_roundtrip(tmp_path):
    from byolo import weights as W
    for variant in W.VARIANTS:
        table = W.layer_table(variant)
        assert len(table) == 75 and table[-1][3] == W.det_channels(variant, 2)
        assert sum(t[5] for t in table) == (15 if variant == 'epistemic' else 0)      # dropout-bearing convs
    w = W.synthetic('aleatoric', 2, 3)
    back = W.unpack(W.pack('aleatoric', w))
    for a, b in zip(w, back):
        assert a.keys() == b.keys() and all(np.array_equal(a[k], b[k]) for k in a)
    n_params = sum(v.size for l in w for v in l.values())
    assert abs(n_params - 61.6e6) < 0.3e6                                           # 61.5 M parameters
    # darknet53.conv.74-style file: backbone only, consumed exactly (darknet.py:66)
    table = W.layer_table('aleatoric')
    path = str(tmp_path / 'dn.conv.74')
    W.write_darknet(path, w[:52], table[:52])
    got = W.read_darknet(path, table[:52])
    assert len(got) == 52 and all(np.array_equal(got[i]['kernel'], w[i]['kernel']) for i in range(52))
    with open(path, 'ab') as f:
        f.write(b'\0\0\0\0')
    with pytest.raises(AssertionError):
        W.read_darknet(path, table[:52])


def _cfg(**kw):
    from lib_yolo import yolov3
    c = {'full_img_size': [96, 160, 3], 'crop': False, 'cls_cnt': 2, 'priors': yolov3.ECP_9_PRIORS, 'aleatoric_loss': False,
         'inference_mode': True, 'T': 4, 'weights': 'synthetic:1', 'implicit_background_class': True}
    c.update(kw)
    return c


def test_model_classes_keep_the_reference_surface():
    from byolo import compat
    from lib_yolo import yolov3
    for cls, oi, ci in ((yolov3.yolov3, 4, 5), (yolov3.yolov3_aleatoric, 9, 11), (yolov3.bayesian_yolov3_aleatoric, 14, 17)):
        y = cls(_cfg())
        assert (y.obj_idx, y.cls_start_idx, y.cls_cnt) == (oi, ci, 2) and list(y.img_size) == [96, 160, 3]
        with pytest.raises(AssertionError):
            y.get_model()
        m = y.init_model(inputs=compat.Placeholder((1, 96, 160, 3)), training=False).get_model()
        assert [(d.h, d.w, d.downsample) for d in m.det_layers] == [(3, 5, 32), (6, 10, 16), (12, 20, 8)]
        assert m.matches_blueprint(y.blueprint) and (m.obj_idx, m.cls_start_idx) == (oi, ci)
        with pytest.raises(Exception, match='only be initialized once'):
            y.init_model(inputs=compat.Placeholder((1, 96, 160, 3)), training=False)
    with pytest.raises(AssertionError):
        yolov3.yolov3(_cfg(full_img_size=[100, 160, 3]))
    with pytest.raises(KeyError):
        yolov3.bayesian_yolov3_aleatoric({k: v for k, v in _cfg().items() if k != 'inference_mode'})
    with pytest.raises(KeyError):
        yolov3.bayesian_yolov3_aleatoric({k: v for k, v in _cfg().items() if k != 'T'})
    # crop rescales priors without touching the shared table (the reference mutates it, model.py:11-15)
    before = yolov3.ECP_9_PRIORS[32][0].h
    y = yolov3.yolov3(_cfg(crop=True, crop_img_size=[64, 96, 3], full_img_size=[128, 192, 3]))
    assert yolov3.ECP_9_PRIORS[32][0].h == before and abs(y.blueprint.det_layers[0].priors[0].h - before * 2) < 1e-12


def test_ecp_records_and_script_surface(tmp_path):
    import detect
    import inference_aleatoric
    import inference_epistemic
    import inference_standard_yolov3
    from byolo import compat, ecp
    from lib_yolo import yolov3
    for mod in (inference_standard_yolov3, inference_aleatoric, inference_epistemic):
        for name in ('Inference', 'concat_bbox', 'nms', 'bbox_to_ecp_format', 'inference', 'main'):
            assert hasattr(mod, name)
    for name in ('box_op_standard', 'box_op_aleatoric', 'box_op_bayes', 'filter_boxes', 'preproces_boxes', 'draw_boxes',
                 'load_img', 'load_model', 'do_it', 'main'):
        assert hasattr(detect, name)
    y = yolov3.bayesian_yolov3_aleatoric(_cfg())
    m = y.init_model(inputs=compat.Placeholder((1, 96, 160, 3)), training=False).get_model()
    row = np.arange(23, dtype=np.float32) / 23
    rec = inference_epistemic.bbox_to_ecp_format(row, [96, 160, 3], m, _cfg())
    assert rec['y0'] == float(row[0] * 96) and rec['x1'] == float(row[3] * 160)
    assert rec['score'] == float(row[14]) * float(row[18]) and rec['identity'] == 'rider'
    assert rec['obj_mutual_info'] == float(row[15]) and rec['cls_entropy'] == float(row[20]) and rec['prior_id'] == float(row[22])
    ya = yolov3.yolov3_aleatoric(_cfg())
    ma = ya.init_model(inputs=compat.Placeholder((1, 96, 160, 3)), training=False).get_model()
    ra = inference_aleatoric.bbox_to_ecp_format(row[:16], [96, 160, 3], ma, _cfg())
    assert ra['cls_entropy'] == ra['layer_id'] == ra['prior_id'] == float(row[13])     # reference quirk kept (:174-176)
    out = ecp.write_ecp_json('epistemic', str(tmp_path), np.stack([row, row]), 'a/b/img_7.png', [96, 160, 3], m, _cfg())
    assert os.path.basename(out) == 'img_7.json' and len(json.load(open(out))['children']) == 2
    # filter / preprocess helpers of detect.py
    boxes = np.stack([row, row * 0.1])
    kept = detect.filter_boxes(boxes, 14, 0.1)
    assert len(kept) == 1
    pp = detect.preproces_boxes([96, 160, 3], kept, 14, 17, 2, _cfg(), {1: 'ped', 2: 'rider'})
    assert pp[0]['cls'] == 'rider' and 0 <= pp[0]['y0'] <= 96
    # weight lookup by step, like the checkpoint lookup of the reference
    os.makedirs(tmp_path / 'ck' / 'run')
    for s in (10, 200):
        open(tmp_path / 'ck' / 'run' / ('weights-%d.byw' % s), 'wb').close()
    c = {'checkpoint_path': str(tmp_path / 'ck'), 'run_id': 'run', 'step': 'last'}
    assert ecp.find_weights(c)[1] == '200' and ecp.find_weights(dict(c, step=10))[1] == '10'


def test_image_dataset_iterates_once(tmp_path):
    from byolo import compat
    rng = np.random.default_rng(0)
    for i in range(3):
        np.save(tmp_path / ('im%d.npy' % i), rng.random((96, 160, 3), dtype=np.float32))
    ds = compat.ImageDataset({'data': {'file_pattern': str(tmp_path / '*.npy')}, 'batch_size': 2, 'full_img_size': [96, 160, 3]})
    src, names = ds.iterator.get_next()
    a = src.next_batch()
    assert a.shape == (2, 96, 160, 3) and src.last_files[0][0].decode().endswith('im0.npy')
    assert src.next_batch().shape[0] == 1
    with pytest.raises(compat.OutOfRangeError):
        src.next_batch()


def _gloo_worker(rank, world, port, n_images, ret):
    import torch.distributed as dist
    sys.path[:0] = [os.path.join(ROOT, 'bayesian-yolov3_b200')]
    from byolo import dist as bd
    dist.init_process_group('gloo', init_method='tcp://127.0.0.1:%d' % port, rank=rank, world_size=world)
    max_out, D = 5, 3

    def run_packed(imgs, index0, out):               # stub hot path: the packed block depends only on the GLOBAL image index
        out.zero_()
        for i in range(imgs.shape[0]):
            g = index0 + i
            cnt = g % 5 + 1
            out[i, :cnt] = float(g) + imgs[i].sum()
            out[i, max_out, 0] = cnt

    imgs = torch.arange(n_images, dtype=torch.float32).view(n_images, 1, 1, 1).expand(n_images, 2, 2, 3).contiguous()
    sd = bd.ShardedDetector(run_packed, n_images, D, max_out=max_out)
    single = torch.zeros((n_images, max_out + 1, D))
    run_packed(imgs, 0, single)
    ok = True
    for slot in (0, 1, 0):                            # slots are reusable; results are views of the slot's receive buffer
        got = sd.result(sd.submit(imgs, slot))
        ok = ok and torch.equal(got.boxes, single[:, :max_out]) and torch.equal(got.counts, single[:, max_out, 0])
        ok = ok and got.counts_int().dtype == torch.int32 and got.boxes.shape == (n_images, max_out, D)
    ret[rank] = bool(ok)
    dist.destroy_process_group()


@pytest.mark.parametrize('n_images', [4, 5])
def test_sharded_detect_equals_single_rank_gloo(n_images):
    import torch.multiprocessing as mp
    from byolo import dist as bd
    assert [bd.shard_range(5, r, 2) for r in range(2)] == [(0, 3), (3, 5)]
    assert [bd.shard_range(32, r, 8) for r in range(8)][-1] == (28, 32)
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 2000 + n_images
    mp.spawn(_gloo_worker, args=(2, port, n_images, ret), nprocs=2, join=True)
    assert ret[0] and ret[1]


def test_det_maps_from_rows_follow_the_reference_dict():
    """vis_uncertainty.py:80-131 reads det_layers[i].det[key] (layers.py:397-411): shapes [lh, lw, 3(, ...)] and the
    values of the row columns they came from."""
    from lib_yolo.model import det_maps_from_rows
    lh, lw, C = 3, 5, 2
    rng = np.random.default_rng(3)
    rows = rng.random((3 * lh * lw, 21 + C)).astype(np.float32)
    det = det_maps_from_rows(rows, lh, lw, C)
    assert det['obj_mean'].shape == (lh, lw, 3) and det['cls_mean'].shape == (lh, lw, 3, C)
    assert det['epi_covar_loc'].shape == (lh, lw, 3, 4, 4) and det['ale_var_loc'].shape == (lh, lw, 3, 4)
    p, y, x = 2, 1, 4
    row = rows[p * lh * lw + y * lw + x]                       # concat_bbox order inside a scale: prior, row, col
    assert det['obj_mean'][y, x, p] == row[14] and det['obj_mutual_info'][y, x, p] == row[15] and det['obj_entropy'][y, x, p] == row[16]
    assert np.array_equal(det['cls_mean'][y, x, p], row[17:19]) and det['cls_mutual_info'][y, x, p] == row[19]
    assert det['cls_entropy'][y, x, p] == row[20]
    assert np.array_equal(np.diagonal(det['epi_covar_loc'][y, x, p]), row[4:8]) and np.array_equal(det['ale_var_loc'][y, x, p], row[8:12])
    assert np.isnan(det['epi_covar_loc'][y, x, p][0, 1])


# ------------------------------------------------------------------------------------------------ TF checkpoints (8f-1)
def test_tf_variable_names_match_the_reference_graph_code():
    """byolo.tf_checkpoint.variable_names vs the names recorded while the reference's own model-building code ran on
    the TF stand-in (tests/golden/tf_variable_names.json, written by tests/golden/gen_golden.py)."""
    import json
    from byolo import tf_checkpoint as TC
    ref = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'tf_variable_names.json')))
    for variant in ('standard', 'aleatoric', 'epistemic'):
        mine = [[n, list(sh)] for n, sh in TC.variable_names(variant, 2)]
        assert mine == ref[variant], variant
    names = [n for n, _ in TC.variable_names('epistemic', 2)]
    assert 'darknet53/conv_46/conv2d/kernel' in names and 'darknet53/downsample_4/batch_normalization/moving_variance' in names
    assert 'det_net_3/detection/conv2d/bias' in names and 'det_net_2/conv_6/conv2d/kernel' in names


def _varint_bytes(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _pb(field, wt, payload):
    tag = _varint_bytes((field << 3) | wt)
    if wt == 0:
        return tag + _varint_bytes(payload)
    return tag + _varint_bytes(len(payload)) + payload


def _table_block(pairs, restart_interval=16):
    out, restarts, last = bytearray(), [], b''
    for i, (k, v) in enumerate(pairs):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(k), len(last)) and k[shared] == last[shared]:
                shared += 1
        out += _varint_bytes(shared) + _varint_bytes(len(k) - shared) + _varint_bytes(len(v)) + k[shared:] + v
        last = k
    for r in restarts or [0]:
        out += np.uint32(r).tobytes()
    out += np.uint32(len(restarts) or 1).tobytes()
    return bytes(out)


def _write_bundle(prefix, tensors, block_entries=40):
    """Test-side writer of the tensor-bundle layout byolo.tf_checkpoint.read_bundle restates (uncompressed blocks)."""
    data, entries = bytearray(), []
    for name in sorted(tensors):
        a = np.asarray(tensors[name], np.float32)          # (ascontiguousarray would turn scalars into shape (1,))
        shape = b''.join(_pb(2, 2, _pb(1, 0, int(d))) for d in a.shape)
        entry = _pb(1, 0, 1) + _pb(2, 2, shape) + _pb(4, 0, len(data)) + _pb(5, 0, a.nbytes)      # dtype DT_FLOAT, shard 0
        entries.append((name.encode(), entry))
        data += a.tobytes()
    pairs = [(b'', _pb(1, 0, 1))] + entries                  # header: num_shards = 1 (little endian = default)
    out, index = bytearray(), []
    for i in range(0, len(pairs), block_entries):
        blk = _table_block(pairs[i:i + block_entries])
        index.append((pairs[min(i + block_entries, len(pairs)) - 1][0], _varint_bytes(len(out)) + _varint_bytes(len(blk))))
        out += blk + b'\0' + b'\0\0\0\0'                    # no compression, crc not checked by the reader
    meta = _table_block([])
    meta_handle = _varint_bytes(len(out)) + _varint_bytes(len(meta))
    out += meta + b'\0' + b'\0\0\0\0'
    iblk = _table_block(index, restart_interval=1)
    index_handle = _varint_bytes(len(out)) + _varint_bytes(len(iblk))
    out += iblk + b'\0' + b'\0\0\0\0'
    footer = meta_handle + index_handle
    out += footer + b'\0' * (40 - len(footer)) + np.uint64(0xdb4775248b80fb57).tobytes()
    open(prefix + '.index', 'wb').write(bytes(out))
    open(prefix + '.data-00000-of-00001', 'wb').write(bytes(data))


def test_tf_checkpoint_round_trip(tmp_path):
    """weights -> variables under the reference's names -> bundle files -> read_bundle -> weights: identical; extra
    variables (optimizer slots, global_step) are ignored; a missing variable is an error."""
    from byolo import tf_checkpoint as TC, weights as W
    ref = W.synthetic('aleatoric', 2, 3)
    variables = TC.variables_from_weights('aleatoric', 2, ref)
    # keep the files small: only variables up to 40k elements are written
    keep = {n: a for n, a in variables.items() if a.size <= 40000}
    keep['darknet53/conv/conv2d/kernel/Adam'] = np.zeros((3, 3, 3, 32), np.float32)       # optimizer slot
    keep['global_step_like'] = np.array(7.0, np.float32)
    prefix = str(tmp_path / 'model.ckpt-42')
    _write_bundle(prefix, keep)
    got = TC.read_bundle(prefix)
    assert set(got) == set(keep)
    for n in keep:
        assert got[n].shape == np.asarray(keep[n]).shape and np.array_equal(got[n], keep[n]), n
    with pytest.raises(KeyError):
        TC.weights_from_variables('aleatoric', 2, got)                                    # big kernels were left out
    full = dict(variables)
    back = TC.weights_from_variables('aleatoric', 2, {k + ':0': v for k, v in full.items()})
    for a, b in zip(back, ref):
        assert set(a) == set(b) and all(np.array_equal(a[k], b[k]) for k in a)
    (tmp_path / 'checkpoint').write_text('model_checkpoint_path: "model.ckpt-42"\n')
    assert TC.latest_checkpoint(str(tmp_path)) == prefix


def test_snappy_block_decoder():
    from byolo.tf_checkpoint import _snappy_decompress
    # literal "abcd", copy (offset 4, length 8) with a 1-byte offset tag, literal "xy"
    src = bytes([14, (4 - 1) << 2]) + b'abcd' + bytes([((8 - 4) << 2) | 1, 4]) + bytes([(2 - 1) << 2]) + b'xy'
    assert _snappy_decompress(src) == b'abcdabcdabcdxy'


def test_find_weights_picks_up_reference_checkpoints(tmp_path):
    """inference_*.py / detect.py lookup (inference_epistemic.py:27-38): BYW1 blobs first, else TF checkpoints by step."""
    from byolo import ecp
    run = tmp_path / 'run1'
    run.mkdir()
    for step in (100, 250):
        for ext in ('.index', '.meta', '.data-00000-of-00001'):
            (run / ('model.ckpt-%d%s' % (step, ext))).write_bytes(b'')
    (run / 'checkpoint').write_text('model_checkpoint_path: "model.ckpt-250"\nall_model_checkpoint_paths: "model.ckpt-100"\n')
    cfg = {'checkpoint_path': str(tmp_path), 'run_id': 'run1', 'step': 'last'}
    assert ecp.find_weights(cfg) == ('tf:' + str(run / 'model.ckpt-250'), '250')
    cfg['step'] = 100
    assert ecp.find_weights(cfg) == ('tf:' + str(run / 'model.ckpt-100'), '100')
    (run / 'weights-7.byw').write_bytes(b'x')
    cfg['step'] = 'last'
    assert ecp.find_weights(cfg) == (str(run / 'weights-7.byw'), '7')


def test_bench_reference_arm_prints_one_json_record():
    """bench.py --impl reference: exactly one stdout line, a JSON record with the contract's keys (the driver parses it)."""
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1'],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    rec = json.loads(lines[0])
    for k in ('impl', 'metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
              'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e'):
        assert k in rec, k
    assert rec['impl'] == 'reference' and rec['value'] > 0 and rec['cpu_baseline']['kind'] == 'port' and rec['cpu_baseline']['cores'] >= 1
    assert rec['e2e']['h2d_bytes_per_step'] == 0 and 'workload' in rec['config'] and 'model' not in rec['config']


def test_ecp_records_have_the_reference_keys_in_the_reference_order():
    """bbox_to_ecp_format of the three inference scripts: same keys in the same order as the reference's dict literals
    (tests/golden/ecp_keys.json, parsed from the reference sources by tests/golden/gen_ecp_keys.py), so json.dump writes
    the files byte for byte like the reference."""
    import json
    from byolo import ecp
    want = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'ecp_keys.json')))

    class M:
        pass
    for variant, (oi, cs, D) in {'standard': (4, 5, 7), 'aleatoric': (9, 11, 16), 'epistemic': (14, 17, 23)}.items():
        m = M()
        m.obj_idx, m.cls_start_idx, m.cls_cnt = oi, cs, 2
        rec = ecp.bbox_to_ecp_format(variant, np.arange(D, dtype=np.float32) / D, (608, 608, 3), m, {'implicit_background_class': True})
        assert list(rec) == want[variant], variant
        assert rec['identity'] == 'rider' and abs(rec['score'] - (oi / D) * ((cs + 1) / D)) < 1e-6


def test_gathered_views_even_and_uneven_shards():
    """byolo.dist.Gathered: the receive buffer [world, per, max_out+1, D] as boxes / counts in global image order; even
    shards are pure views of the buffer, uneven shards drop the padding images of the short ranks."""
    from byolo import dist as bd
    world, per, max_out, D = 3, 2, 4, 5
    recv = torch.arange(world * per * (max_out + 1) * D, dtype=torch.float32).view(world, per, max_out + 1, D)
    even = bd.Gathered(recv, world * per, world, max_out)
    assert even.boxes.shape == (6, max_out, D) and even.counts.shape == (6,)
    assert even.boxes.data_ptr() == recv.data_ptr()                         # a view, not a copy
    assert torch.equal(even.boxes[3], recv[1, 1, :max_out]) and even.counts[3] == recv[1, 1, max_out, 0]
    uneven = bd.Gathered(recv, 5, world, max_out)                           # shards 2, 2, 1
    assert uneven.boxes.shape == (5, max_out, D)
    assert torch.equal(uneven.boxes[4], recv[2, 0, :max_out]) and torch.equal(uneven.boxes[3], recv[1, 1, :max_out])
    assert [bd.shard_range(5, r, 3) for r in range(3)] == [(0, 2), (2, 4), (4, 5)]


def test_bench_arms_share_one_config_dict():
    """The reference arm and the B200 arm of bench.py must print the same `config` (the driver compares them) and the
    default configuration is the one BASELINE.json quotes the metric on (configs[2])."""
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location('bench_mod', os.path.join(ROOT, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    base = json.load(open(os.path.join(ROOT, 'BASELINE.json')))
    assert sorted(bench.CONFIGS) == [1, 2, 3, 4, 5] and len(base['configs']) == 5
    c3 = bench.CONFIGS[3]
    assert (c3['variant'], c3['T'], c3['batch_per_gpu'], c3['img']) == ('epistemic', 10, 16, 608)
    assert (bench.CONFIGS[4]['T'], bench.CONFIGS[4]['img'], bench.CONFIGS[4]['batch_per_gpu']) == (30, 416, 4)
    assert bench.CONFIGS[5]['variant'] == 'nms-stress' and bench.CONFIGS[5]['batch_per_gpu'] == 8
    for n, c in bench.CONFIGS.items():
        assert c['workload'].startswith('configs[%d]' % (n - 1)) and 'model' not in c
