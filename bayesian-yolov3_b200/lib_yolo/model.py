"""Containers the reference's scripts read from a built model (reference: lib_yolo/model.py:188-268).
The graph builder itself (ModelBuilder, model.py:20-185) is replaced by the execution plan inside libbyolo."""
import numpy as np

from byolo.priors import Prior


def img_size_and_priors_if_crop(config):
    """model.py:6-17.  Returns (img_size, priors); priors are rescaled when cropping.  Unlike the reference this does
    not write the rescaled priors back into config['priors'] (the reference mutates the shared module-level table)."""
    img_size = config['crop_img_size'] if config.get('crop') else config['full_img_size']
    priors = config['priors']
    if config.get('crop'):
        sh = config['full_img_size'][0] / float(config['crop_img_size'][0])
        sw = config['full_img_size'][1] / float(config['crop_img_size'][1])
        priors = {s: [Prior(h=p.h * sh, w=p.w * sw) for p in prs] for s, prs in priors.items()}
    return img_size, priors


class DetLayerBlueprint:
    def __init__(self, input_img_size, downsample_factor, priors):
        self.h = input_img_size[0] // downsample_factor
        self.w = input_img_size[1] // downsample_factor
        self.downsample = downsample_factor
        self.priors = priors


class ModelBlueprint:
    def __init__(self, det_layers, cls_cnt):
        self.det_layers = det_layers
        self.cls_cnt = cls_cnt


class _Activations:
    """Lazy read-back of conv outputs of the most recent run (byolo_get_activation): nothing is copied off the device
    unless a script actually looks at `raw_output`, `dn_out`, `det_net_*_out` or `layers[i]`."""

    def __init__(self):
        self.engine, self.run_id, self.cache = None, 0, {}

    def bind(self, engine):
        self.engine, self.run_id, self.cache = engine, self.run_id + 1, {}

    def get(self, conv_index):
        if self.engine is None:
            return None                                           # no run yet (graph tensors have no value either)
        if conv_index not in self.cache:
            self.cache[conv_index] = self.engine.activation(conv_index).cpu().numpy()
        return self.cache[conv_index]


class _LayerList:
    """`Model.layers`: the outputs of the 75 convolutions in creation (= weight file) order, read lazily.  The
    reference's list (model.py:40-41) also holds the residual adds, routes and upsamplings; here the add is part of the
    block-closing conv's output and routes / upsamplings are not materialised."""

    def __init__(self, acts):
        self._acts = acts

    def __len__(self):
        return 75

    def __getitem__(self, i):
        return self._acts.get(range(75)[i])


DN_OUT_CONV = 51                       # darknet53 output (reference layer 74)
DET_CONVS = (58, 66, 74)               # the three detection convs (raw head outputs)


class DetLayer(DetLayerBlueprint):
    """One detection scale.  `bbox` (list of 3 per-prior arrays [..,g,g,D]), `raw_output` and `det` hold the values of
    the most recent run (None before the first run) - in the reference they are graph tensors."""

    def __init__(self, input_img_size, downsample_factor, priors, layer_id, acts=None):
        super().__init__(input_img_size, downsample_factor, priors)
        self.layer_id = layer_id
        self.loc_loss = self.obj_loss = self.cls_loss = None
        self.bbox = self.det = None
        self._acts = acts

    @property
    def raw_output(self):
        """Output of the detection conv (model.py:122,151,184): [B (or T), g, g, 3*(5+C) | 3*2*(5+C)] fp32."""
        return self._acts.get(DET_CONVS[self.layer_id]) if self._acts is not None else None

    def matches_blueprint(self, bp):
        return (self.h, self.w, self.downsample, len(self.priors)) == (bp.h, bp.w, bp.downsample, len(bp.priors)) and all(
            p.h == q.h and p.w == q.w for p, q in zip(self.priors, bp.priors))


class Model:
    """What `yolo.init_model(...).get_model()` returns: detection layers + column indices + an executor."""

    def __init__(self, variant, engine_factory, inputs, img_size, priors, cls_cnt, obj_idx, cls_start_idx, T):
        self.variant, self.inputs, self.cls_cnt = variant, inputs, cls_cnt
        self.obj_idx, self.cls_start_idx, self.T = obj_idx, cls_start_idx, T
        self.img_size = img_size
        self._acts = _Activations()
        self.det_layers = [DetLayer(img_size, s, priors[s], i, self._acts) for i, s in enumerate((32, 16, 8))]
        self.layers = _LayerList(self._acts)
        self._engine_factory, self._engine, self._max_batch = engine_factory, None, 0

    # yolov3.py:306-310, 624-628: backbone output and the raw outputs of the three detection sub-nets
    dn_out = property(lambda self: self._acts.get(DN_OUT_CONV))
    det_net_1_out = property(lambda self: self._acts.get(DET_CONVS[0]))
    det_net_2_out = property(lambda self: self._acts.get(DET_CONVS[1]))
    det_net_3_out = property(lambda self: self._acts.get(DET_CONVS[2]))

    def matches_blueprint(self, blueprint):
        return self.cls_cnt == blueprint.cls_cnt and all(
            dl.matches_blueprint(bp) for dl, bp in zip(self.det_layers, blueprint.det_layers))

    def engine(self, batch):
        if self._engine is None or batch > self._max_batch:
            if self._engine is not None:
                self._engine.close()
            self._engine, self._max_batch = self._engine_factory(batch), batch
        return self._engine

    def execute(self, img, seed=0, image_index0=0, max_out=1000):
        """One pass of the hot path for a host batch [B,H,W,3]: returns dict(rows, boxes, count, idx) of numpy arrays
        and refreshes det_layers[*].bbox like a fetch of those tensors would."""
        import torch
        img = np.ascontiguousarray(img, np.float32)
        assert img.ndim == 4 and tuple(img.shape[1:]) == tuple(self.img_size), (img.shape, self.img_size)
        if self.variant == 'epistemic':
            assert img.shape[0] == 1, 'the Bayesian model processes one image per run (inference_epistemic.py:193)'
        eng = self.engine(img.shape[0])
        boxes, cnt, idx, rows = eng.detect(torch.from_numpy(img).to(eng.device), seed=seed, image_index0=image_index0,
                                           max_out=max_out, want_rows=True)
        torch.cuda.synchronize(eng.device)
        self._acts.bind(eng)
        res = dict(rows=rows.cpu().numpy(), boxes=boxes.cpu().numpy(), count=cnt.cpu().numpy(), idx=idx.cpu().numpy())
        off = 0
        for dl in self.det_layers:                              # per-prior views in the reference's shapes
            n = dl.h * dl.w
            per = [res['rows'][:, off + p * n: off + (p + 1) * n].reshape(-1, dl.h, dl.w, res['rows'].shape[-1]) for p in range(3)]
            dl.bbox = [p[0] for p in per] if self.variant == 'epistemic' else per
            if self.variant == 'epistemic':
                dl.det = det_maps_from_rows(res['rows'][0, off: off + 3 * n], dl.h, dl.w, self.cls_cnt)
            off += 3 * n
        return res


def det_maps_from_rows(rows, lh, lw, cls_cnt):
    """The per-cell statistics dict of `decode_epistemic` (reference layers.py:397-411) that vis_uncertainty.py:80-131
    colour-maps, rebuilt from the decoded rows of ONE detection scale ([3*lh*lw, 21+C], prior-major as concat_bbox
    writes them).  Shapes follow the reference: [lh, lw, 3(, 4 | 4,4 | C)].  The kernel exports the diagonal of the
    epistemic covariance only (the reference's consumers index [..., i, i]); off-diagonal entries are NaN.  `ev_loc`,
    `obj_samples` and `cls_samples` (marked irrelevant in the reference) are not exported."""
    C = cls_cnt
    r = np.asarray(rows).reshape(3, lh, lw, -1).transpose(1, 2, 0, 3)          # [lh, lw, prior, D]
    cov = np.full(r.shape[:3] + (4, 4), np.nan, r.dtype)
    for i in range(4):
        cov[..., i, i] = r[..., 4 + i]
    return {
        'epi_covar_loc': cov,
        'ale_var_loc': r[..., 8:12],
        'obj_mean': r[..., 14], 'obj_mutual_info': r[..., 15], 'obj_entropy': r[..., 16],
        'cls_mean': r[..., 17:17 + C], 'cls_mutual_info': r[..., 17 + C], 'cls_entropy': r[..., 18 + C],
    }
