"""Shared implementation of the three inference scripts (reference: inference_standard_yolov3.py, inference_aleatoric.py,
inference_epistemic.py - three near-identical files there as well).  `surface(variant)` returns the names each script
exports: Inference(yolo, config).run(), concat_bbox, nms, bbox_to_ecp_format, inference(config), main().  One JSON
file per image in ECP format; the forward pass, decode and NMS run in libbyolo (sm_100a), not in a TensorFlow session.
"""
import json
import logging
import os
import threading
import time

from byolo import compat as tf          # Session / errors.OutOfRangeError stand-ins
from byolo import ecp

# per variant: model class name in lib_yolo.yolov3, default run id, and the config keys that differ in the reference's main()
VARIANTS = {
    'standard': dict(cls='yolov3', run_id='yolov3', extra={'batch_size': 11, 'inference_mode': False}),
    'aleatoric': dict(cls='yolov3_aleatoric', run_id='ale', extra={'batch_size': 11, 'inference_mode': False}),
    # batch 1 only: decode_epistemic reduces the whole batch axis (inference_epistemic.py:193); T: "edit if OOM errors"
    'epistemic': dict(cls='bayesian_yolov3_aleatoric', run_id='epi_ale', extra={'batch_size': 1, 'T': 50, 'inference_mode': True}),
}


def surface(variant):
    from lib_yolo import dataset_utils, yolov3
    spec = VARIANTS[variant]

    class Inference:
        def __init__(self, yolo, config):
            self.batch_size = config['batch_size']
            dataset = dataset_utils.TestingDataset(config)
            self.img_tensor, self.filename_tensor = dataset.iterator.get_next()
            weights, step = ecp.find_weights(config)
            self.img_size = config['full_img_size']
            assert not config['crop']
            self.out_path = '{}_{}'.format(config['out_path'], step)
            os.makedirs(self.out_path)
            self.config = config
            self.worker_thread = None
            if variant == 'epistemic':
                assert config['inference_mode']
            yolo.load_weights(weights)
            self.model = yolo.init_model(inputs=self.img_tensor, training=False).get_model()
            bbox = concat_bbox([dl for dl in self.model.det_layers], model=self.model)
            self.nms = nms(bbox, self.model, per_class=config.get('per_class_nms', False))

        def run(self):
            with tf.Session(seed=self.config.get('seed', 0)) as sess:
                self.sess = sess
                step = 0
                while True:
                    try:
                        step += 1
                        processed = self.process_batch(sess)
                        if variant != 'epistemic' or step % 15 == 0:
                            logging.info('Processed {} images.'.format((step - 1) * self.batch_size + processed))
                    except tf.errors.OutOfRangeError:
                        break
                if self.worker_thread:
                    self.worker_thread.join()
            return self

        def process_batch(self, sess):
            boxes, files = sess.run([self.nms, self.filename_tensor])
            counts = sess.last_counts
            if self.worker_thread:
                self.worker_thread.join()
            if variant == 'epistemic':
                boxes = boxes[None]                                   # [n,23] -> [1,n,23]
                counts = [boxes.shape[1]]
            # the arrays are caller-owned copies: the worker serialises them while the next batch runs
            self.worker_thread = threading.Thread(target=self.write_to_disc, args=(boxes, files, counts))
            self.worker_thread.start()
            return len(files)

        epistemic_forward_pass = process_batch                        # name used by inference_epistemic.py:75

        def write_to_disc(self, all_boxes, files, counts):
            for b, filename in enumerate(files):
                self.write_ecp_json(all_boxes[b][:counts[b]], filename[0].decode('utf-8'))

        def write_ecp_json(self, boxes, img_name):
            return ecp.write_ecp_json(variant, self.out_path, boxes, img_name, self.img_size, self.model, self.config)

    def concat_bbox(net_out, model=None):
        """Fetch handle for all candidate rows in (scale, prior, row, col) order.  net_out: the three detection layers (or
        their .bbox lists, as the reference passes them); the rows are produced at their final offsets by the decode
        kernel, so nothing is concatenated here."""
        if model is None:
            model = getattr(net_out[0], 'model', None)
        assert model is not None, 'pass model= when handing over plain bbox lists'
        return tf.RowsOp(model)

    def nms(boxes, model, per_class=False):
        """Class-agnostic NMS (max 1000 boxes, IoU 0.5, no score threshold) + gather, as a fetch handle.  per_class=True
        (config['per_class_nms']): the reference's commented variant "used to produce the results for the paper"
        (inference_epistemic.py:104-126): one NMS per class over the rows where that class scores highest, concatenated."""
        return tf.NmsOp(boxes, model, per_class=per_class)

    def bbox_to_ecp_format(bbox, img_size, model, config):
        return ecp.bbox_to_ecp_format(variant, bbox, img_size, model, config)

    def inference(config):
        if variant == 'epistemic':
            assert config['batch_size'] == 1
        assert not config['crop']
        logging.info(json.dumps(config, indent=4, default=lambda x: str(x)))
        logging.info('----- START -----')
        start = time.time()
        yolo = getattr(yolov3, spec['cls'])(config)
        Inference(yolo, config).run()
        elapsed = int(time.time() - start)
        logging.info('----- FINISHED in {:02d}:{:02d}:{:02d} -----'.format(elapsed // 3600, (elapsed // 60) % 60, elapsed % 60))

    def main():
        config = {
            'checkpoint_path': './checkpoints',  # edit
            'run_id': spec['run_id'],  # edit
            'step': 'last',  # edit: int or 'last'
            'full_img_size': [1024, 1920, 3],  # edit if not ECP dataset
            'cls_cnt': 2,  # edit if not ECP dataset
            'cpu_thread_cnt': 24,
            'crop': False,
            'training': False,
            'aleatoric_loss': False,
            'priors': yolov3.ECP_9_PRIORS,  # edit
            'implicit_background_class': True,
            'data': {
                'path': '$HOME/data/ecp/images',  # edit: image files instead of the reference's tfrecords
                'file_pattern': '*.png',  # edit
            },
        }
        config.update(spec['extra'])
        config['data']['file_pattern'] = os.path.join(os.path.expandvars(config['data']['path']), config['data']['file_pattern'])
        config['out_path'] = os.path.join('./inference', config['run_id'])  # edit
        inference(config)

    return dict(VARIANT=variant, Inference=Inference, concat_bbox=concat_bbox, nms=nms, bbox_to_ecp_format=bbox_to_ecp_format,
                inference=inference, main=main)


def run_as_script(main):
    import numpy as np
    np.set_printoptions(suppress=True, formatter={'float_kind': '{:5.3}'.format})
    logging.basicConfig(level=logging.DEBUG, format='%(asctime)s, pid: %(process)d, %(levelname)-8s %(message)s',
                        datefmt='%a, %d %b %Y %H:%M:%S')
    main()
