"""The three model classes of the reference behind their original constructor / init_model / get_model surface
(reference: lib_yolo/yolov3.py:176-628), executing on libbyolo instead of building a TensorFlow graph.

Config keys honoured (same meaning as in the reference): full_img_size, crop, crop_img_size, cls_cnt, priors,
aleatoric_loss (accepted, training only), inference_mode, T, standard_test_dropout, freeze_darknet53 (accepted).
Extra optional keys: 'precision' ('fp16' tensor-core path | 'fp32' exact CUDA-core path), 'weights' (list of 75 layer
dicts, a BYW1 blob, a path to one, or 'synthetic:<seed>'), 'seed' (dropout stream).
"""
import os

from byolo import priors as _priors
from byolo import weights as _weights
from lib_yolo import model

CITY_PERSONS_9_PRIORS = _priors.by_stride('CITY_PERSONS_9_PRIORS')
ECP_9_PRIORS = _priors.by_stride('ECP_9_PRIORS')
ECP_NIGHT_9_PRIORS = _priors.by_stride('ECP_NIGHT_9_PRIORS')
ECP_DAY_NIGHT_9_PRIORS = _priors.by_stride('ECP_DAY_NIGHT_9_PRIORS')
ECP_BIC_9_PRIORS = _priors.by_stride('ECP_BIC_9_PRIORS')


class _Base:
    variant = None
    obj_idx = cls_start_idx = None

    def __init__(self, config):
        self._model = None
        self._config = config
        self.img_size, self._priors = model.img_size_and_priors_if_crop(config)
        self.cls_cnt = config['cls_cnt']
        self._T = 1
        self.blueprint = model.ModelBlueprint(
            det_layers=[model.DetLayerBlueprint(self.img_size, s, self._priors[s]) for s in (32, 16, 8)], cls_cnt=self.cls_cnt)
        # input size must be a multiple of the largest stride (yolov3.py:207-211)
        assert config['full_img_size'][0] % 32 == 0
        assert config['full_img_size'][1] % 32 == 0
        if config.get('crop'):
            assert config['crop_img_size'][0] % 32 == 0
            assert config['crop_img_size'][1] % 32 == 0
        self._weights = None

    def get_model(self):
        """call init_model first!"""
        assert self._model is not None, 'Call init_model first.'
        return self._model

    def _resolve_weights(self):
        if self._weights is not None:
            return self._weights
        w = self._config.get('weights', None)
        if w is None:
            raise RuntimeError("no weights: set config['weights'] (layer list, BYW1 blob/path, 'tf:<checkpoint prefix>' or 'synthetic:<seed>'), "
                               "or call load_darknet53_weights / load_weights")
        if isinstance(w, str) and w.startswith('synthetic'):
            seed = int(w.split(':')[1]) if ':' in w else 0
            w = _weights.synthetic(self.variant, self.cls_cnt, seed)
        elif isinstance(w, str) and w.startswith('tf:'):          # a TensorFlow checkpoint prefix of the reference model
            from byolo import tf_checkpoint
            w = tf_checkpoint.load(w[3:], self.variant, self.cls_cnt)
        elif isinstance(w, str):
            with open(os.path.expandvars(w), 'rb') as f:
                w = f.read()
        self._weights = w
        return w

    def load_weights(self, weights):
        """Replaces tf.train.Saver().restore: `weights` as for config['weights']."""
        self._config['weights'] = weights
        self._weights = None
        if self._model is not None and self._model._engine is not None:
            self._model._engine.load_weights(self._resolve_weights())

    def load_darknet53_weights(self, weightfile):
        """darknet.load_darknet_weights (darknet.py:42-122): overwrites the 52 backbone convs from a darknet53.conv.74-style
        file; the head keeps the weights that are configured (whatever form they were given in: layer list, BYW1 blob or
        path, TF checkpoint, 'synthetic:<seed>').  Without configured weights the head is synthetic seed 0 - said aloud."""
        assert self._model is not None, 'Call init_model first.'
        table = _weights.layer_table(self.variant, self.cls_cnt)
        if self._weights is None and self._config.get('weights') is None:
            import logging
            logging.warning('load_darknet53_weights: no head weights configured, using synthetic(seed 0) for the head')
            base = _weights.synthetic(self.variant, self.cls_cnt, 0)
        else:
            base = self._resolve_weights()
            if isinstance(base, (bytes, bytearray)):
                base = _weights.unpack(bytes(base))
        base = list(base)
        _weights.read_darknet(weightfile, table[:52], into=base)
        self.load_weights(base)
        return []                                             # the reference returns assign ops for sess.run

    def init_model(self, inputs, training, gt1=None, gt2=None, gt3=None):
        if self._model is not None:
            raise Exception('model can only be initialized once!')
        assert not training and gt1 is None and gt2 is None and gt3 is None, 'training graphs are out of scope'
        in_shape = list(inputs.get_shape())
        assert len(in_shape) == 4, 'invalid data format'
        cfg = self._config

        def factory(batch):
            import byolo
            eng = byolo.Engine(self.variant, self.img_size[:2], self.cls_cnt, priors=self._priors, T=self._T,
                               max_batch=batch, precision=cfg.get('precision', 'fp16'),
                               standard_test_dropout=cfg.get('standard_test_dropout', False))
            return eng.load_weights(self._resolve_weights())

        self._model = model.Model(self.variant, factory, inputs, tuple(self.img_size), self._priors, self.cls_cnt,
                                  self.obj_idx, self.cls_start_idx, self._T)
        assert self._model.matches_blueprint(self.blueprint), 'Model does not match blueprint'
        return self


class yolov3(_Base):                                          # yolov3.py:176
    variant, obj_idx, cls_start_idx = 'standard', 4, 5


class yolov3_aleatoric(_Base):                                # yolov3.py:313
    variant, obj_idx, cls_start_idx = 'aleatoric', 9, 11

    def __init__(self, config):
        config['aleatoric_loss']                              # required key, as in the reference (yolov3.py:315)
        super().__init__(config)


class bayesian_yolov3_aleatoric(_Base):                       # yolov3.py:454
    obj_idx, cls_start_idx = 14, 17

    def __init__(self, config):
        config['aleatoric_loss']
        self._inference_mode = config['inference_mode']       # required key (yolov3.py:461)
        super().__init__(config)
        if self._inference_mode:
            self._T = config['T']                             # required in inference mode (yolov3.py:467-468)

    variant = 'epistemic'

    def init_model(self, inputs, training, gt1=None, gt2=None, gt3=None):
        if not self._inference_mode:
            # inference_mode=False is the training/validation graph of the reference (dropout on, no MC stacking,
            # aleatoric decode, model.py:169-170); it is used by uncertainty_training.py only.
            raise NotImplementedError('bayesian_yolov3_aleatoric(inference_mode=False) is the training graph: out of scope')
        return super().init_model(inputs, training, gt1, gt2, gt3)
