"""Single-image detection with the same call surface as the reference's detect.py: box_op_*, filter_boxes,
preproces_boxes, draw_boxes, load_img, load_model(sess, config, model_cls), do_it(...), main().
The model runs in libbyolo; `sess` is the byolo.compat.Session stand-in."""
import glob
import logging
import os

import numpy as np

import inference_aleatoric
import inference_epistemic
import inference_standard_yolov3
from byolo import compat as tf
from byolo import ecp
from lib_yolo import yolov3


def box_op_standard(model):
    bbox = inference_standard_yolov3.concat_bbox(model.det_layers, model=model)
    return inference_standard_yolov3.nms(bbox, model)[0, ...]


def box_op_aleatoric(model):
    bbox = inference_aleatoric.concat_bbox(model.det_layers, model=model)
    return inference_aleatoric.nms(bbox, model)[0, ...]


def box_op_bayes(model):
    bbox = inference_epistemic.concat_bbox(model.det_layers, model=model)
    return inference_epistemic.nms(bbox, model)


def filter_boxes(boxes, obj_idx, thresh):
    """Rows whose objectness exceeds `thresh` (detect.py:36-37), as a list of rows like the reference returns."""
    boxes = np.asarray(boxes)
    if boxes.size == 0:
        return []
    return list(boxes[boxes[:, obj_idx] > thresh])


def preproces_boxes(img_size, boxes, obj_idx, cls_start_idx, cls_cnt, config, cls_mapping=None):
    out = []
    for box in boxes:
        cls_col = int(np.argmax(box[cls_start_idx:cls_start_idx + cls_cnt]))
        cls_idx = cls_col + 1 if config['implicit_background_class'] else cls_col
        cls = cls_mapping[cls_idx] if cls_mapping else cls_idx
        # FLAGGED deviation: the reference indexes the score AFTER the +1 background shift (detect.py:44-51), which
        # reads the neighbouring column (or raises IndexError for the 7-column standard rows); the winning class's
        # own score is used here.
        cls_score = box[cls_col + cls_start_idx]
        y0, x0, y1, x1 = (np.clip(box[i], 0, 1) * img_size[i % 2] for i in range(4))
        out.append({'cls': cls, 'score': box[obj_idx] * cls_score, 'obj_score': box[obj_idx], 'cls_score': cls_score,
                    'y0': y0, 'x0': x0, 'y1': y1, 'x1': x1})
    return out


def draw_boxes(img, boxes, color=(43, 219, 216), thickness=1):
    """Overlay of label, score and rectangle on a float RGB image in [0,1] (out of the hot path; same signature and look
    as the reference helper, detect.py:66-73)."""
    import cv2
    rgb = tuple(float(c) / 255.0 for c in color)
    for b in boxes:
        top_left, bottom_right = (int(b['x0']), int(b['y0'])), (int(b['x1']), int(b['y1']))
        cv2.rectangle(img, top_left, bottom_right, rgb, thickness)
        cv2.putText(img, '%s %4.3f' % (b['cls'], b['score']), top_left, cv2.FONT_HERSHEY_SIMPLEX, 0.5, rgb, thickness)


def _centre_crop(img, size_hw):
    """The window of size_hw around the image centre (what detect.py:79-82 cuts out when config['crop'] is set)."""
    off = [(full - want) // 2 for full, want in zip(img.shape[:2], size_hw[:2])]
    return img[off[0]:off[0] + size_hw[0], off[1]:off[1] + size_hw[1]]


def load_img(config, img_size, filename):
    """float32 RGB in [0,1], centre-cropped when config['crop'], with a leading batch axis (detect.py:76-85)."""
    if filename.endswith('.npy'):
        img = np.load(filename).astype(np.float32)
    else:
        import cv2
        img = cv2.imread(filename, cv2.IMREAD_COLOR)[:, :, ::-1].astype(np.float32) / np.float32(255.0)
    if config['crop']:
        img = _centre_crop(img, img_size)
    return np.ascontiguousarray(img)[None]


def load_model(sess, config, model_cls):
    if model_cls == yolov3.bayesian_yolov3_aleatoric:
        config['inference_mode'] = True
    yolo = model_cls(config)
    img_tensor = tf.Placeholder(shape=(1, *yolo.img_size))
    weights, _ = ecp.find_weights(config)
    yolo.load_weights(weights)
    model = yolo.init_model(inputs=img_tensor, training=False).get_model()
    return model, img_tensor


def do_it(files, thresh, config, model_cls, cls_mapping, show=False):
    box_op = {yolov3.yolov3: box_op_standard, yolov3.yolov3_aleatoric: box_op_aleatoric,
              yolov3.bayesian_yolov3_aleatoric: box_op_bayes}[model_cls]
    results = {}
    with tf.Session(seed=config.get('seed', 0)) as sess:
        model, img_tensor = load_model(sess, config, model_cls)
        img_size = img_tensor.shape.as_list()[1:]
        op = box_op(model)                                    # built once (the reference re-adds it per file)
        for file in files:
            img = load_img(config, img_size, file)
            boxes, = sess.run([op], feed_dict={img_tensor: img})
            if model.variant != 'epistemic':
                boxes = boxes[:sess.last_counts[0]]
            boxes = filter_boxes(boxes, model.obj_idx, thresh)
            boxes = preproces_boxes(img_size, boxes, model.obj_idx, model.cls_start_idx, model.cls_cnt, config,
                                    cls_mapping=cls_mapping)
            img = img[0, ...]
            draw_boxes(img, boxes)
            logging.info('{}: {}'.format(os.path.basename(file), boxes))
            results[file] = boxes
            if show:                                          # interactive display is out of scope; save instead
                import cv2
                os.makedirs(config['out_path'], exist_ok=True)
                cv2.imwrite(os.path.join(config['out_path'], os.path.basename(file) + '.det.png'),
                            (np.clip(img[:, :, ::-1], 0, 1) * 255).astype(np.uint8))
    return results


def main():
    config = {
        'checkpoint_path': './checkpoints/',
        'run_id': 'epi_ale',  # edit
        'step': 'last',  # edit: int or 'last'
        'crop_img_size': [768, 1440, 3],
        'full_img_size': [1024, 1920, 3],  # edit if not ecp
        'cls_cnt': 2,  # edit if not ecp
        'T': 35,  # only relevant for bayesian model
        'cpu_thread_cnt': 10,
        'freeze_darknet53': False,
        'crop': False,
        'training': False,
        'aleatoric_loss': True,
        'priors': yolov3.ECP_9_PRIORS,
        'out_path': './uncertainty_visualization',  # edit
        'implicit_background_class': True,  # whether the label ids start at 1 or 0. True = 1, False = 0
    }
    cls_mapping = {1: 'ped', 2: 'rider'}  # edit; {0: 'ped', 1: 'rider'} if labels start at 0
    thresh = 0.1  # edit
    files = glob.glob('./test_images/*')  # edit
    model_cls = yolov3.bayesian_yolov3_aleatoric  # edit: yolov3.yolov3 | yolov3.yolov3_aleatoric
    do_it(files, thresh, config, model_cls, cls_mapping, show=True)


if __name__ == '__main__':
    logging.basicConfig(level=logging.DEBUG, format='%(asctime)s, pid: %(process)d, %(levelname)-8s %(message)s',
                        datefmt='%a, %d %b %Y %H:%M:%S')
    main()
