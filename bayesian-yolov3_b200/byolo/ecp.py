"""ECP-format JSON records and the threaded writer shared by the three inference scripts
(reference: bbox_to_ecp_format + write_ecp_json in inference_standard_yolov3.py:87-169, inference_aleatoric.py:85-178,
inference_epistemic.py:85-170)."""
import json
import os

import numpy as np

LABEL_TO_CLS_NAME = {1: 'pedestrian', 2: 'rider'}            # ECP; starts at 0 without implicit background class

# (json key, column) tables in the reference's key order; 'o+k' = obj_idx + k, 'c+k' = cls_start_idx + cls_cnt + k, 'score' and
# 'cls_scores' are the computed entries.  The aleatoric table reproduces the reference as written: its cls_entropy, layer_id
# and prior_id all read column cls_start_idx + cls_cnt (inference_aleatoric.py:174-176).  Same keys, same order, same
# values: json.dump writes the files byte for byte as the reference does.
_COLUMNS = {
    'standard': ['score', 'cls_scores'],
    'aleatoric': [('x_var', 4), ('y_var', 5), ('w_var', 6), ('h_var', 7), ('total_var', 8), 'score', ('obj_entropy', 'o+1'),
                  'cls_scores', ('cls_entropy', 'c+0'), ('layer_id', 'c+0'), ('prior_id', 'c+0')],
    'epistemic': [('x_var_epi', 4), ('y_var_epi', 5), ('w_var_epi', 6), ('h_var_epi', 7), ('x_var_ale', 8), ('y_var_ale', 9),
                  ('w_var_ale', 10), ('h_var_ale', 11), ('total_var_epi', 12), ('total_var_ale', 13), 'score',
                  ('obj_mutual_info', 'o+1'), ('obj_entropy', 'o+2'), 'cls_scores', ('ped_score', 17), ('rider_score', 18),
                  ('cls_mutual_info', 'c+0'), ('cls_entropy', 'c+1'), ('layer_id', 'c+2'), ('prior_id', 'c+3')],
}


def bbox_to_ecp_format(variant, bbox, img_size, model, config):
    h, w = img_size[:2]
    cs, cc, oi = model.cls_start_idx, model.cls_cnt, model.obj_idx
    cls_scores = bbox[cs:cs + cc]
    cls_idx = int(np.argmax(cls_scores))
    label = cls_idx + 1 if config['implicit_background_class'] else cls_idx
    rec = {'y0': float(bbox[0] * h), 'x0': float(bbox[1] * w), 'y1': float(bbox[2] * h), 'x1': float(bbox[3] * w)}
    for entry in _COLUMNS[variant]:
        if entry == 'score':
            rec['score'] = float(bbox[oi]) * float(bbox[cs + cls_idx])
        elif entry == 'cls_scores':
            rec['cls_scores'] = cls_scores
        else:
            key, col = entry
            if isinstance(col, str):
                col = (oi if col[0] == 'o' else cs + cc) + int(col[2:])
            rec[key] = float(bbox[col])
    rec['identity'] = LABEL_TO_CLS_NAME.get(label, label)
    return rec


def write_ecp_json(variant, out_path, boxes, img_name, img_size, model, config):
    out_file = os.path.join(out_path, '{}.json'.format(os.path.splitext(os.path.basename(img_name))[0]))
    with open(out_file, 'w') as f:
        json.dump({'children': [bbox_to_ecp_format(variant, b, img_size, model, config) for b in boxes]}, f,
                  default=lambda x: x.tolist())
    return out_file


def find_weights(config):
    """Checkpoint lookup of the inference scripts (reference: inference_epistemic.py:27-38, detect.py:96-105):
    config['checkpoint_path']/config['run_id'] holds either BYW1 blobs named weights-<step>.byw or the reference's own
    TensorFlow checkpoints (`step == 'last'`: tf.train.latest_checkpoint via the `checkpoint` state file, else the prefix
    whose `-<step>.meta` / `-<step>.index` exists).  config['weights'] overrides the lookup.
    Returns (weights argument for load_weights, step label); a TF checkpoint comes back as 'tf:<prefix>'."""
    if config.get('weights') is not None:
        return config['weights'], str(config.get('step', 'given'))
    folder = os.path.join(config['checkpoint_path'], config['run_id'])
    blobs = {}
    for f in os.listdir(folder):
        if f.startswith('weights-') and f.endswith('.byw'):
            blobs[int(f[len('weights-'):-len('.byw')])] = os.path.join(folder, f)
    if blobs:
        step = max(blobs) if config['step'] == 'last' else int(config['step'])
        assert step in blobs, 'could not find checkpoint'
        return blobs[step], str(step)
    from . import tf_checkpoint
    if config['step'] == 'last':
        prefix = tf_checkpoint.latest_checkpoint(folder)
    else:
        prefix = None
        for f in sorted(os.listdir(folder)):
            if f.endswith('-{}.meta'.format(config['step'])) or f.endswith('-{}.index'.format(config['step'])):
                prefix = os.path.join(folder, os.path.splitext(f)[0])
                break
    assert prefix is not None, 'could not find checkpoint'
    return 'tf:' + prefix, prefix.split('-')[-1]
